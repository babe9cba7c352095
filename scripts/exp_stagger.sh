#!/bin/bash
# T = 5000 (R = 5, 320 threads, two CTAs per SM fit): does staggering two co-resident CTAs overlap FP64 and shared-memory phases?
mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 3 --frames 5000 --atoms 200000 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], 'step', d['ms_per_step'], d['fft_plan'])"; }
TA_B200_K1_PATH=r16 run r16
TA_B200_K1_PATH=r8 TA_B200_K1E_ONE_CTA=1 run r8_one_cta
TA_B200_K1_PATH=r8 run r8_two_ctas
for st in 1500 3000 6000 12000 24000; do TA_B200_K1_PATH=r8 TA_B200_K1E_STAGGER=$st run r8_stagger_$st; done

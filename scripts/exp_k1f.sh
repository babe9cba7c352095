#!/bin/bash
# K1 fast-path experiments on the GPU box: plain / NB=2 / per-phase clock profile
run() { python bench.py --steps 4 --warmup 3 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], 'e2e_ms', d['e2e']['ms_per_step'], d['fft_plan'])"; }
run plain
TA_B200_K1F_NB=2 run nb2
TA_B200_K1F_PREFETCH=1 run prefetch
TA_B200_K1F_NB=2 TA_B200_K1F_PREFETCH=1 run nb2_prefetch
TA_B200_K1F_PROFILE=1 run profiled; grep -A40 "k1f profile" gpurun_out/exp_err_profiled.txt | tail -34

#!/bin/bash
# K1 fast-path experiments on the GPU box
run() { python bench.py --steps 4 --warmup 3 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], d['fft_plan'])"; }
for st in 0 1500 3000 5000 8000 12000 20000 40000 80000; do TA_B200_K1F_STAGGER=$st run stagger$st; done
for st in 3000 8000 20000; do TA_B200_K1F_NT=160 TA_B200_K1F_STAGGER=$st run nt160_stagger$st; done

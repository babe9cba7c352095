#!/bin/bash
# K1 fast-path timing experiments (wrong results by construction): which pipe bounds the kernel?
run() { python bench.py --steps 4 --warmup 3 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], d['fft_plan'])"; }
for v in NOFP NOSMEM; do TA_B200_LIB=$PWD/build/exp_$v.so run $v; done

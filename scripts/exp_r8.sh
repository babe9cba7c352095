#!/bin/bash
# radix-8 K1 path: parity, then one bench line per path
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "radix8 or fast_path" 2>&1 | tail -15
run() { python bench.py --steps 4 --warmup 3 "${@:2}" 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['fft_plan'])"; }
TA_B200_K1_PATH=r8 run r8
TA_B200_K1_PATH=r16 run r16

#!/bin/bash
run() { python bench.py --steps 4 --warmup 3 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['fft_plan'])"; }
run smemtables

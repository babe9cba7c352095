#!/bin/bash
# K1 three-pass kernel experiments: one bench line per (frames, atoms, TA_B200_K1F_VAR) triple, e.g. "10000:100000:4"
mkdir -p gpurun_out
export TA_B200_K1_PATH=r16
run() { python bench.py --steps 4 --warmup 3 --frames $2 --atoms $3 2>gpurun_out/exp_err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'k1_ms', d['roofline']['kernel_ms'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['fft_plan'])"; }
for spec in "$@"; do
  IFS=: read T A v <<< "$spec"
  TA_B200_K1F_VAR=$v python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fast_path_vs_oracle" 2>&1 | tail -1
  TA_B200_K1F_VAR=$v run T${T}_var$v $T $A
done

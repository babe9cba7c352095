#!/bin/bash
# windowed / Helfand kernels on the GPU box
for wl in windowed helfand helfand_direct; do
python bench.py --workload $wl --steps 5 --warmup 3 2>gpurun_out/exp_err_$wl.txt | tee gpurun_out/bench_$wl.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', 'kernel_ms', r['kernel_ms'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'fp64 frac', r['fp64']['frac'], 'af/s', d['value'])"
done

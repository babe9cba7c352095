#!/bin/bash
# K1 at the headline size with the three series layouts: float series + particle sums in shared memory (default for float
# sources), float series + global particle sums, double series (TA_B200_SERIES = narrow_gpart / wide)
mkdir -p gpurun_out
for v in ${@:-default narrow_gpart wide}; do
  TA_B200_SERIES=$v timeout 300 python bench.py --workload fft --steps 5 --warmup 3 > gpurun_out/k1s_$v.json 2> gpurun_out/k1s_$v.err
  python - "$v" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/k1s_{n}.json'))
    print(n, 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'parity', d.get('parity_check'), 'e2e', round(d['e2e']['ms_per_step'],1), d['fft_plan'])
except Exception as e:
    print(n, 'FAILED', e); print(open(f'gpurun_out/k1s_{n}.err').read()[-1500:])
PY
done

#!/bin/bash
# K1 at the headline size in the builds that exist side by side.  Prints kernel_ms of each.
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload fft --steps 5 --warmup 3 > gpurun_out/k1v_$name.json 2> gpurun_out/k1v_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/k1v_{n}.json'))
    print(n, 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'parity', d.get('parity_check'), 'e2e', round(d['e2e']['ms_per_step'],1))
except Exception as e:
    print(n, 'FAILED', e); print(open(f'gpurun_out/k1v_{n}.err').read()[-1500:])
PY
}
for v in "$@"; do
  case $v in
    pipe) run pipe A=1 ;;
    pipe_s*) run $v TA_B200_K1P_STAGGER=${v#pipe_s} ;;
    pref_s*) run $v TA_B200_K1P_PREF=1 TA_B200_K1P_STAGGER=${v#pref_s} ;;
    pipe_pref) run pipe_pref TA_B200_K1P_PREF=1 ;;
    lockstep) run lockstep TA_B200_K1_PATH=lockstep ;;
    wide) run wide TA_B200_K1_PATH=lockstep TA_B200_SERIES=wide ;;
  esac
done

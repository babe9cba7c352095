#!/bin/bash
# one bench line per "NAME=VALUE[,NAME=VALUE...]" environment setting (use "-" for the defaults); default workload
mkdir -p gpurun_out
for spec in "$@"; do
  tag=$(echo "$spec" | tr -c 'A-Za-z0-9\n' '_')
  ( if [ "$spec" != "-" ]; then IFS=, read -ra kv <<< "$spec"; for e in "${kv[@]}"; do export "$e"; done; fi
    python bench.py --workload fft --steps 4 --warmup 3 2>gpurun_out/exp_err_$tag.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$spec', 'k1_ms', d['roofline']['kernel_ms'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['fft_plan'])" )
done

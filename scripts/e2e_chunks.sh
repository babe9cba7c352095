#!/bin/bash
# end-to-end rate of the bulk staging path against the particle-chunk size (row length of the 2-D H2D copies), N ranks at once
N=${1:-4}; shift
mkdir -p gpurun_out
for c in "$@"; do
  TA_B200_BULK_CHUNK=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --workload fft --steps 4 --warmup 2 > gpurun_out/e2e_chunk_${N}_$c.json 2> gpurun_out/e2e_chunk_${N}_$c.err
  python - $N $c <<'PY'
import json,sys
n,c=sys.argv[1:]
try:
    d=json.load(open(f'gpurun_out/e2e_chunk_{n}_{c}.json'))
    print('ranks',n,'chunk',c,'e2e ms',round(d['e2e']['ms_per_step'],1),'GB/s per gpu',round(d['e2e']['h2d_gbs_per_gpu'],1),'probe',[round(x,1) for x in d['h2d_probe']['per_rank_gbs']],'dev ms',round(d['ms_per_step'],2))
except Exception as e:
    print('FAILED',n,c,e); print(open(f'gpurun_out/e2e_chunk_{n}_{c}.err').read()[-800:])
PY
done

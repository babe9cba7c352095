#!/bin/bash
# Round-end profiling on one B200 (run under gpurun): the launch list of the bench command and one --set full capture of
# the whole-shard K1 launch of the device-resident leg.  Outputs under gpurun_out/; summaries are made on the build box.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fft.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/launches_fft_bench.log 2>&1
# 20,000 atoms: the end-to-end leg launches K1 once per staging chunk, the device-resident leg once per step; the last K1
# launch of the run is a whole-shard launch -> count them first, then capture that one
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k1 --csv --log-file gpurun_out/k1_launches_20k.csv \
    python bench.py --atoms 20000 --steps 1 --warmup 1 > /dev/null 2>&1
n=$(grep -c "k1" gpurun_out/k1_launches_20k.csv)
echo "K1 launches at 20,000 atoms: $n"
ncu --set full --clock-control none --import-source on -k regex:k1 -s $((n - 1)) -c 1 -f -o gpurun_out/k1_full \
    python bench.py --atoms 20000 --steps 1 --warmup 1 > gpurun_out/ncu_k1_full.log 2>&1
tail -2 gpurun_out/ncu_k1_full.log | cut -c1-200

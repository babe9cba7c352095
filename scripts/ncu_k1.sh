#!/bin/bash
# one --set full capture of the K1 kernel selected by TA_B200_K1_PATH ($1), 20,000 atoms x 10,000 frames
mkdir -p gpurun_out
export TA_B200_K1_PATH=$1
ncu --set full --clock-control none --import-source on -k regex:k1 -s 2 -c 1 -f -o gpurun_out/k1_$1 \
    python bench.py --atoms 20000 --steps 1 --warmup 1 > gpurun_out/ncu_k1_$1.log 2>&1
tail -3 gpurun_out/ncu_k1_$1.log

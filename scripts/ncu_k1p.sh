#!/bin/bash
# one --set full capture of the whole-shard launch of the pipelined K1 (20,000 atoms); env passes through
name=${1:-k1p}
A="--workload fft --atoms 20000 --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k1p_fft_acf --csv --log-file gpurun_out/cnt.csv python bench.py $A > /dev/null 2>&1
n=$(grep -c k1p_fft_acf gpurun_out/cnt.csv); echo "launches $n"
ncu --set full --clock-control none --import-source on -k regex:k1p_fft_acf -s $((n - 1)) -c 1 -f -o gpurun_out/r02_$name python bench.py $A > gpurun_out/r02_ncu_$name.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_$name.ncu-rep gpurun_out/r02_${name}_ncu_summary.json --command "ncu --set full --clock-control none -k regex:k1p_fft_acf -s $((n-1)) -c 1 python bench.py $A" > /dev/null 2>&1
rm -f gpurun_out/cnt.csv

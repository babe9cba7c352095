#!/bin/bash
mkdir -p gpurun_out
A="--workload fft --atoms 20000 --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k1f_fft_acf --csv --log-file gpurun_out/cnt.csv python bench.py $A > /dev/null 2>&1
n=$(grep -c k1f_fft_acf gpurun_out/cnt.csv); echo "K1 launches $n"
ncu --set full --clock-control none --import-source on -k regex:k1f_fft_acf -s $((n - 1)) -c 1 -f -o gpurun_out/r02_k1_fp64_tmem python bench.py $A > gpurun_out/r02_ncu_k1_fp64_tmem.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_k1_fp64_tmem.ncu-rep gpurun_out/r02_k1_fp64_tmem_ncu_summary.json --command "ncu --set full --clock-control none -k regex:k1f_fft_acf -s $((n-1)) -c 1 python bench.py $A" > /dev/null 2>&1
rm -f gpurun_out/cnt.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_all_v2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench_v2.log 2>&1
bash scripts/sanitize.sh > gpurun_out/r02_compute_sanitizer_v2.txt 2>&1; grep -c "0 errors\|0 hazards" gpurun_out/r02_compute_sanitizer_v2.txt; grep -v "0 errors\|0 hazards" gpurun_out/r02_compute_sanitizer_v2.txt | head -5

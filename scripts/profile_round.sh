#!/bin/bash
# Round profiling on one B200 (run under gpurun): the launch list of the default bench command and one --set full capture
# of each shipped kernel.  Outputs under gpurun_out/; scripts/ncu_summary.py turns the reports into profiles/*.json here.
mkdir -p gpurun_out
R=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_all.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/${R}_launches_bench.log 2>&1
cap() {  # name, kernel regex, launches to skip, bench arguments...
  name=$1; rx=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${R}_$name \
      python bench.py "$@" > gpurun_out/${R}_ncu_$name.log 2>&1
  tail -1 gpurun_out/${R}_ncu_$name.log | cut -c1-160
  # the summary is made here (gpurun brings back at most 64 MiB); only the reports named in KEEP travel
  python scripts/ncu_summary.py gpurun_out/${R}_$name.ncu-rep gpurun_out/${R}_${name}_ncu_summary.json \
      --command "ncu --set full --clock-control none -k regex:$rx -s $skip -c 1 python bench.py $*" > /dev/null 2>&1
  case " $KEEP " in *" $name "*) ;; *) rm -f gpurun_out/${R}_$name.ncu-rep ;; esac
}
KEEP=${KEEP:-"k1_fp64"}
# K1: the e2e leg launches it once per staging chunk, the device-resident leg once per step over the whole shard: the
# LAST K1 launch of a run is a whole-shard launch -> count them first
count() { ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$1 --csv --log-file gpurun_out/${R}_cnt.csv \
          python bench.py "${@:2}" > /dev/null 2>&1; grep -c "$1" gpurun_out/${R}_cnt.csv; }
A="--workload fft --atoms 20000 --steps 1 --warmup 1"
n=$(count k1f_fft_acf $A); echo "K1 fp64 launches: $n"; cap k1_fp64 k1f_fft_acf $((n - 1)) $A
A="--workload fft_fp32 --atoms 20000 --steps 1 --warmup 1"
n=$(count k1f_fft_acf $A); echo "K1 fp32 launches: $n"; cap k1_fp32 k1f_fft_acf $((n - 1)) $A
A="--workload windowed --steps 1 --warmup 1"
n=$(count k_windowed $A); echo "K2 launches: $n"; cap k2 k_windowed $((n - 1)) $A
A="--workload helfand --steps 1 --warmup 1"
n=$(count k_windowed $A); echo "K3 launches: $n"; cap k3 k_windowed $((n - 1)) $A
A="--workload helfand_fft --atoms 60000 --frames 10000 --steps 1 --warmup 1"
n=$(count k5_helfand $A); echo "K5 launches: $n"; cap k5 k5_helfand $((n - 1)) $A; cap k6 k6_helfand $((n - 1)) $A
# K0: any chunk launch of the headline staging (2,072 particles x 10,000 frames)
cap k0 k0_stage 3 --workload fft --atoms 20000 --steps 1 --warmup 1
rm -f gpurun_out/${R}_cnt.csv; ls -la gpurun_out/

#!/bin/bash
# round-end measurements on one B200: GPU test-suite, one bench line per workload, the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fft.json 2>gpurun_out/bench_fft.err; cut -c1-400 gpurun_out/bench_fft.json
for w in windowed helfand helfand_direct; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; cut -c1-300 gpurun_out/bench_$w.json
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
python __graft_entry__.py --smoke 2>&1 | tail -2

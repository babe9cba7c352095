#!/bin/bash
# round-end measurements on one B200: GPU test-suite, the all-workload bench line, the reference arm, smoke()
mkdir -p gpurun_out
R=${1:-r02}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_all.json 2>gpurun_out/${R}_bench_all.err; cut -c1-400 gpurun_out/${R}_bench_all.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>gpurun_out/${R}_bench_reference.err; cut -c1-400 gpurun_out/${R}_bench_reference.json
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2

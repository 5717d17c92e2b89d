#!/bin/bash
# one ncu --set full capture of two mid-circuit sweep launches of the benchmark
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 25 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out/ | grep prof

#!/bin/bash
# GPU parity tests, one ncu --set full capture of two mid-circuit sweep launches, an un-profiled bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 3 gpurun_out/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>gpurun_out/bench_quick.err
tail -c 600 gpurun_out/bench_quick.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 28 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out/ | grep prof

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "pipelined or wb12" > gpurun_out/pytest_e2e.log 2>&1; tail -n 3 gpurun_out/pytest_e2e.log
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.log 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
tail -c 900 gpurun_out/bench_e2e.log; tail -n 5 gpurun_out/bench_e2e.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --gpus 1 --qubits 33 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_1gpu_33q.log 2> gpurun_out/bench_1gpu_33q.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_1gpu_33q.log; tail -n 3 gpurun_out/bench_1gpu_33q.err

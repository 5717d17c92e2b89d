#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 800 -p no:cacheprovider -x -k "specialised or wb28" > gpurun_out/pytest_jit.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_jit.log
bash tools/gpu_bench_matrix.sh "|" "|--steps 5"

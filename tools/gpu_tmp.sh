#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 800 -p no:cacheprovider -x -k "specialised or wb24 or wb28 or round_trip" > gpurun_out/pytest_jit.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_jit.log
bash tools/gpu_bench_matrix.sh "|" "QFB_PLAN_LATE=0|" "QFB_JIT_L2PF=1|" "QFB_JIT_MINB=4|" "|--tile-bits 12" "QFB_JIT_MINB=6|"

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_autograd.py -m gpu -q --timeout 200 -p no:cacheprovider -x -k "batch" > gpurun_out/pytest_autograd.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_autograd.log

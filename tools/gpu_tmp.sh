#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 280 -p no:cacheprovider -x -s -k "wb30_depth3" > gpurun_out/pytest_30q_d3.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_30q_d3.log

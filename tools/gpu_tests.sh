#!/bin/bash
# GPU parity tests (optionally only the files given as arguments)
mkdir -p gpurun_out
timeout 2400 python -m pytest ${@:-tests} -m gpu -q --timeout 1500 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_fuzz_jit.py 100 7 > gpurun_out/r2_fuzz_jit.log 2>&1; echo "fuzz rc=$?"; tail -n 5 gpurun_out/r2_fuzz_jit.log

#!/bin/bash
mkdir -p gpurun_out
QFB_JIT=1 QFB_REG_BITS=4 timeout 600 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 900 -p no:cacheprovider -x 2>&1 | tail -n 40 > gpurun_out/r4_fail.log
QFB_REG_BITS=4 timeout 900 ncu --set full --clock-control none -k regex:qfb_sweep -s 17 -c 2 -f -o gpurun_out/prof_jit_r4 \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_jit_r4.log 2>&1
tail -2 gpurun_out/ncu_jit_r4.log

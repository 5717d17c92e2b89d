#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:qfb_sweep -s 19 -c 2 -f -o gpurun_out/prof_jit_m11 \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --tile-bits 11 > gpurun_out/ncu_jit_m11.log 2>&1
tail -2 gpurun_out/ncu_jit_m11.log
timeout 900 ncu --set full --clock-control none -k regex:qfb_sweep -s 17 -c 2 -f -o gpurun_out/prof_jit_m12a \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_jit_m12a.log 2>&1
tail -2 gpurun_out/ncu_jit_m12a.log

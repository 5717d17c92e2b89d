#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:qfb_sweep -s 19 -c 2 -f -o gpurun_out/prof_jit_final \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_jit_final.log 2>&1
tail -2 gpurun_out/ncu_jit_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
tail -3 gpurun_out/r2_launches.csv

#!/bin/bash
# ncu capture of the sweep-specialised kernels of the benchmark plan (3 launches, full set) + launch list
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qfb_sweep -s 17 -c 3 -f -o gpurun_out/prof_jit \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_jit.log 2>&1
tail -3 gpurun_out/ncu_jit.log
ls -la gpurun_out/prof_jit.ncu-rep

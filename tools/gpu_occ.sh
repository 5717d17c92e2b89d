#!/bin/bash
# occupancy sensitivity of the sweep kernel: pad dynamic shared memory to force fewer CTAs per SM
mkdir -p gpurun_out
for cfg in "12 0" "12 60000" "11 0" "11 9000" "11 20000" "11 40000" "10 0" "10 12000"; do
  set -- $cfg
  QFB_SMEM_PAD=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile-bits $1 2>> gpurun_out/occ.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('OCC tile $1 pad $2', d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" | tee -a gpurun_out/occ.log
done

#!/bin/bash
# tile-order / prefetch-granularity knobs of the sweep kernel on the benchmark
mkdir -p gpurun_out
for cfg in "0 3" "1 3" "2 3" "2 7"; do
  set -- $cfg
  QFB_TILE_ORDER=$1 QFB_PF_MASK=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/knobs.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('KNOB order $1 pfmask $2', d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" | tee -a gpurun_out/knobs.log
done

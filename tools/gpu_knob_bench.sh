#!/bin/bash
# quick bench lines for a few planner knob settings (no e2e / cpu baseline)
mkdir -p gpurun_out
for cfg in "12 3 28" "12 2 28" "12 1 28"; do
  set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile-bits $1 --low-bits $2 --max-cost $3 2>> gpurun_out/sweep_knobs.err | tee -a gpurun_out/bench_knobs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH', d['plan']['tile_bits'], '$2', d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
done

#!/bin/bash
# benchmark under sets of environment settings: each argument is "VAR=val,VAR=val,..."
mkdir -p gpurun_out
for cfg in "$@"; do
env $(echo $cfg | tr ',' ' ') timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/jit.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH $cfg', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
done
tail -3 gpurun_out/jit.err

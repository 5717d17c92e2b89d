#!/bin/bash
# benchmark matrix: each argument is "VAR=val,VAR=val|bench.py arguments" (either side may be empty)
mkdir -p gpurun_out
for cfg in "$@"; do
envs=${cfg%%|*}; args=${cfg#*|}
env $(echo $envs | tr ',' ' ') timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $args 2>> gpurun_out/matrix.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH [$cfg]', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f plan_s %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['plan'].get('plan_seconds',-1)))"
done
tail -3 gpurun_out/matrix.err

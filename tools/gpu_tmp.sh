#!/bin/bash
mkdir -p gpurun_out
for t in 12 11 10 9; do
QFB_TILE_BITS=$t timeout 300 python bench.py --config c1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 tile $t', 'ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step'])"
done
QFB_JIT=1 timeout 300 python bench.py --config c1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 jit', 'ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step'])"
QFB_JIT=1 QFB_TILE_BITS=10 timeout 300 python bench.py --config c1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 jit tile 10', 'ms', round(d['ms_per_step'],3), 'launches', d['launches_per_step'])"

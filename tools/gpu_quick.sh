#!/bin/bash
# GPU parity tests + one quick bench line (no e2e / cpu baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>> gpurun_out/quick.err | tee gpurun_out/bench_quick.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"

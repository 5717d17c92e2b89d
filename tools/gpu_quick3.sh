#!/bin/bash
# GPU parity tests + quick bench lines for tile sizes 12 (default), 13 and 11 (no e2e / cpu baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 3 gpurun_out/pytest.log
: > gpurun_out/bench_quick.log
for tb in 12 13 11; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile-bits $tb 2>> gpurun_out/quick.err | tee -a gpurun_out/bench_quick.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f plan_s %.1f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['plan']['plan_seconds']))"
done

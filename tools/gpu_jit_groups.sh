#!/bin/bash
# sweep-specialised kernels: tiles per CTA (QFB_JIT_GROUPS) -- parity then benchmark per setting
mkdir -p gpurun_out
for g in "$@"; do
QFB_JIT=1 QFB_JIT_GROUPS=$g timeout 600 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 900 -p no:cacheprovider -x 2>&1 | tail -n 2
QFB_JIT_GROUPS=$g timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/jit.err | tee gpurun_out/bench_jit_g$g.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH groups=$g', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
done
tail -5 gpurun_out/jit.err

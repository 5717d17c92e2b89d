#!/bin/bash
# sweep-specialised kernels: async tile copy on/off x register bits x tile bits. args: "A:R:M" triples
mkdir -p gpurun_out
for arm in "$@"; do
IFS=: read a r m <<< "$arm"
QFB_JIT=1 QFB_JIT_ASYNC=$a QFB_REG_BITS=$r timeout 600 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 900 -p no:cacheprovider -x 2>&1 | tail -n 2
QFB_JIT_ASYNC=$a QFB_REG_BITS=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile-bits $m 2>> gpurun_out/jit.err | tee gpurun_out/bench_jit_a${a}_r${r}_m$m.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH async=$a R=$r M=$m', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
done
tail -5 gpurun_out/jit.err

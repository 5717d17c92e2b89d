#!/bin/bash
# sweep-specialised kernels: parity tests with the JIT forced on, then the benchmark with and without it
mkdir -p gpurun_out
QFB_JIT=1 timeout 1200 python -m pytest tests/test_gpu_circuits.py tests/test_gpu_programs.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_jit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_jit.log
tail -n 15 gpurun_out/pytest_jit.log
for jit in 1 0; do
QFB_JIT=$jit timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>> gpurun_out/jit.err | tee gpurun_out/bench_jit$jit.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH jit=$jit', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f plan_s %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['plan'].get('plan_seconds', -1)))"
done
tail -5 gpurun_out/jit.err

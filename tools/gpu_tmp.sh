#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_autograd.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_autograd.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_autograd.log
for i in 1 2; do
timeout 300 python bench.py --config c2 --steps 20 > gpurun_out/r2_bench_c2.json 2> gpurun_out/c2.err; echo "c2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_c2.json').read().strip().splitlines()[-1]); print('fused', d['ms_per_step'], d['launches_per_step'])"
QFB_SMALL_CIRCUIT=0 timeout 300 python bench.py --config c2 --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gate by gate', d['ms_per_step'], d['launches_per_step'])"
done

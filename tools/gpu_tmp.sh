#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['plan']['rounds'], d['e2e']['value'], d['clocks'])
PY

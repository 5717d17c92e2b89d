#!/bin/bash
# the driver's round-end sequence on one GPU: reference arm, default bench line, the other configurations
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/full.err; tail -c 600 gpurun_out/r2_bench_reference.json; echo
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2>> gpurun_out/full.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print('BENCH ms/step %.1f value %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))
for k in ('e2e','e2e_reference_shaped','e2e_cold','cpu_baseline','clocks','plan'):
    print(k, json.dumps(d.get(k))[:400])
PY
for c in c1 c2 c3; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_$c.json 2>> gpurun_out/full.err; cut -c1-700 gpurun_out/r2_bench_$c.json; done
tail -5 gpurun_out/full.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
: > gpurun_out/sweep_knobs.jsonl
for cfg in "12 3 28" "11 3 28" "12 3 12"; do
  set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --tile-bits $1 --low-bits $2 --max-cost $3 >> gpurun_out/sweep_knobs.jsonl 2>> gpurun_out/sweep_knobs.err
done
tail -n 3 gpurun_out/pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/sweep_knobs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))
PY

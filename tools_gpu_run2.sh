#!/bin/bash
# perf exploration: planner knobs on the 30q benchmark + tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
: > gpurun_out/sweep_knobs.jsonl
for cfg in "12 3 28" "12 3 14" "12 3 9" "12 3 6" "12 4 28" "11 3 28" "13 3 28" "13 3 14"; do
  set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --tile-bits $1 --low-bits $2 --max-cost $3 >> gpurun_out/sweep_knobs.jsonl 2>> gpurun_out/sweep_knobs.err
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 25 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
tail -n 5 gpurun_out/pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/sweep_knobs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))
PY

#!/bin/bash
# GPU parity tests, op micro-benchmarks and a quick bench line (no e2e / cpu baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 5 gpurun_out/pytest.log
python tools_microbench.py 30 12 2>&1 | tee gpurun_out/microbench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-62s %7.2f ms rounds %d' % (d['case'][:62], d['ms'], d['rounds']))"
for cfg in "12 3 28" "12 3 40"; do
  set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile-bits $1 --low-bits $2 --max-cost $3 2>> gpurun_out/sweep_knobs.err | tee -a gpurun_out/bench_knobs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print('BENCH', d['plan']['tile_bits'], d['plan']['sweeps'], d['plan']['rounds'], 'ms/step %.1f gates/s %.0f frac %.3f sweep_ms %.2f'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
done

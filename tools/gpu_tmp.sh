#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "matches_single" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bm2.log 2> gpurun_out/bm2.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bm2.log'):
    try: d=json.loads(l)
    except Exception: continue
    print('ms/step %.1f value %.0f sweeps %s parity %s comm %s' % (d['ms_per_step'], d['value'], d.get('plan',{}).get('sweeps'), d.get('parity_max_abs'), {k:v for k,v in d.get('comm',{}).items() if k in ('ms_per_step','pipelined_remaps_per_step','sweeps_inside_pipelines_per_step')}))
PY

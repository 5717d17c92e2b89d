#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "matches_single" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_multi.log
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
run() {
cfg=$1; shift
env $(echo $cfg | tr ',' ' ') $TR 29517 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/bm2.log 2> gpurun_out/bm2.err; echo "rc=$? [$cfg $@]"
python - <<'PY'
import json
for l in open('gpurun_out/bm2.log'):
    try: d=json.loads(l)
    except Exception: continue
    print('ms/step %.1f value %.0f sweeps %s parity %s comm %s' % (d['ms_per_step'], d['value'], d.get('plan',{}).get('sweeps'), d.get('parity_max_abs'), {k:v for k,v in d.get('comm',{}).items() if k in ('ms_per_step','pipelined_remaps_per_step','sweeps_inside_pipelines_per_step')}))
PY
grep -v "CudaIPC\|OMP_NUM\|\*\*\*\*" gpurun_out/bm2.err | tail -n 2
}
run QFB_REMAP_CHAIN=1 --qubits 30
run QFB_REMAP_CHAIN=2 --qubits 30
run QFB_REMAP_CHAIN=3 --qubits 30
run QFB_REMAP_CHAIN=2 --no-parity
run QFB_REMAP_CHAIN=3 --no-parity

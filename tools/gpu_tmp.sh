#!/bin/bash
mkdir -p gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
for cfg in "QFB_REMAP_SLICE_BITS=2" "QFB_REMAP_SLICE_BITS=3" "QFB_REMAP_SLICE_BITS=2,QFB_SLICE_ROOM=0" "QFB_REMAP_SLICE_BITS=2,QFB_REMAP_CTAS=2" "QFB_REMAP_SLICE_BITS=3,QFB_SLICE_ROOM=2,QFB_REMAP_CTAS=2" "QFB_REMAP_SLICE_BITS=1"; do
env $(echo $cfg | tr ',' ' ') $TR 29517 bench.py --gpus 2 --qubits 30 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bm2.log 2> gpurun_out/bm2.err; echo "rc=$? [$cfg]"
python - <<'PY'
import json
for l in open('gpurun_out/bm2.log'):
    try: d=json.loads(l)
    except Exception: continue
    print('ms/step %.1f value %.0f parity %s comm %s' % (d['ms_per_step'], d['value'], d.get('parity_max_abs'), {k:v for k,v in d.get('comm',{}).items() if k in ('ms_per_step','pipelined_remaps_per_step')}))
PY
grep -v "CudaIPC" gpurun_out/bm2.err | tail -n 2
done

#!/bin/bash
# the driver's scaling run for one N: reference arm, then the default bench line (33 qubits per GPU when N > 1)
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_ref_$N.json 2> gpurun_out/scale_$N.err
tail -c 300 gpurun_out/r2_scale_ref_$N.json; echo
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2_scale_$N.json 2>> gpurun_out/scale_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_$N.json').read().strip().splitlines()[-1])
print('SCALE N=$N qubits', d['config']['qubits'], 'ms/step %.1f value %.0f circuit gates/s %.1f sweeps %d frac %.3f parity %s norm_err %.1e'%(d['ms_per_step'], d['value'], d['circuit_gates_per_s'], d['plan']['sweeps'], d['roofline']['frac'], d.get('parity_max_abs'), d['norm_error_after_run']))
print('comm', json.dumps(d.get('comm'))[:300]); print('e2e', json.dumps(d.get('e2e'))[:300]); print('cold', json.dumps(d.get('e2e_cold'))[:200])
PY
tail -3 gpurun_out/scale_$N.err

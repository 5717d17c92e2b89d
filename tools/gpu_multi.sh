#!/bin/bash
# multi-GPU: sharded parity tests, then bench with the peer-memory and the NCCL exchange. arg1 = GPUs, rest = qubits per GPU
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -n 3
for q in "$@"; do
for path in peer nccl; do
QFB_REMAP=$path timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --qubits $q --no-e2e --no-cpu-baseline 2>> gpurun_out/multi.err | tee gpurun_out/bench_multi_${N}_q${q}_$path.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    c=d.get('comm',{})
    print('BENCH N=$N q=$q $path sweeps', d['plan']['sweeps'], 'ms/step %.1f circuit gates/s %.0f norm_err %.1e comm ms %.1f remaps %.1f GB/rank %.1f'%(d['ms_per_step'], d['circuit_gates_per_s'], d['norm_error_after_run'], c.get('ms_per_step',0), c.get('remaps_per_step',0), c.get('bytes_sent_per_rank_per_step',0)/1e9))"
done
done
tail -5 gpurun_out/multi.err

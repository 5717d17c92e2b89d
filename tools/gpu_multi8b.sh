#!/bin/bash
# 8 GPUs: weak-scaling bench at 30 qubits per GPU (33 qubits), no e2e
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_multi_8_q30.log 2> gpurun_out/bench_multi_8_q30.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_multi_8_q30.log; tail -n 3 gpurun_out/bench_multi_8_q30.err

#!/bin/bash
# 2-GPU run: sharded parity tests, then the weak-scaling bench at 30 qubits per GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -n 4 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_multi_2_q30.log 2> gpurun_out/bench_multi_2_q30.err
echo "rc=$?"; tail -c 1200 gpurun_out/bench_multi_2_q30.log; tail -n 3 gpurun_out/bench_multi_2_q30.err

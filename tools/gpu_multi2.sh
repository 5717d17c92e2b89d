#!/bin/bash
# 2-GPU run: sharded parity tests, then the weak-scaling bench at 30 and 33 qubits per GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi_multi.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -n 4 gpurun_out/pytest_multi.log
for q in 30 33; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --qubits $q --no-e2e > gpurun_out/bench_multi_2_q$q.log 2> gpurun_out/bench_multi_2_q$q.err
  echo "rc=$?"; tail -c 1800 gpurun_out/bench_multi_2_q$q.log; tail -n 5 gpurun_out/bench_multi_2_q$q.err
done

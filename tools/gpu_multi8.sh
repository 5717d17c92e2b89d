#!/bin/bash
# 8-GPU run: sharded parity tests (world 2/4/8), weak-scaling bench at 30 and 33 qubits per GPU (33 and 36 qubits)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi_multi8.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
tail -n 4 gpurun_out/pytest_multi8.log
for q in 30 33; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 --qubits $q --no-e2e > gpurun_out/bench_multi_8_q$q.log 2> gpurun_out/bench_multi_8_q$q.err
  echo "rc=$?"; tail -c 1500 gpurun_out/bench_multi_8_q$q.log; tail -n 3 gpurun_out/bench_multi_8_q$q.err
done

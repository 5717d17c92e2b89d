#!/bin/bash
# multi-GPU pass: sharded parity tests + weak-scaling bench at N = this box's GPU count
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_multi_$N.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_multi_$N.log
tail -n 6 gpurun_out/pytest_multi.log; tail -n 3 gpurun_out/bench_multi_$N.log

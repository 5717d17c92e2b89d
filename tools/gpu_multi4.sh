#!/bin/bash
# 4 GPUs: the default bench line as the driver runs it (with e2e)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_multi_4_default.log 2> gpurun_out/bench_multi_4_default.err; echo "rc=$?"
tail -c 1400 gpurun_out/bench_multi_4_default.log; tail -n 3 gpurun_out/bench_multi_4_default.err

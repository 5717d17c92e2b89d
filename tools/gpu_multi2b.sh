#!/bin/bash
# 2 GPUs: exchange bandwidth under a few NCCL settings, then the default bench line (with e2e) as the driver runs it
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
$TR tools/exchange_bench.py 30 2>&1 | grep EXCHANGE | tee gpurun_out/exchange_bench.log
NCCL_NCHANNELS_PER_PEER=32 $TR tools/exchange_bench.py 30 2>&1 | grep EXCHANGE | tee -a gpurun_out/exchange_bench.log
NCCL_MIN_NCHANNELS=32 NCCL_NCHANNELS_PER_PEER=32 $TR tools/exchange_bench.py 30 2>&1 | grep EXCHANGE | tee -a gpurun_out/exchange_bench.log
timeout 1500 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_multi_2_default.log 2> gpurun_out/bench_multi_2_default.err; echo "rc=$?"
tail -c 1200 gpurun_out/bench_multi_2_default.log; tail -n 3 gpurun_out/bench_multi_2_default.err

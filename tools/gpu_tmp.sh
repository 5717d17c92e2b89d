#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 > gpurun_out/r2_scale_2.json 2> gpurun_out/scale_2.err; echo "rc=$?"
tail -c 3000 gpurun_out/r2_scale_2.json | cut -c1-1500; grep -v "CudaIPC\|OMP_NUM\|\*\*\*\*" gpurun_out/scale_2.err | tail -n 5

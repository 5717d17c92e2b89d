#!/usr/bin/env python
"""Times the in-place remap exchange of ShardedCircuit alone (one rank bit, 2^nl-amplitude shards) for a few staging
sizes; run under torchrun. NCCL knobs come from the environment."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantumflow_b200 import sharded      # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
shard = torch.zeros(1 << nl, dtype=torch.complex128, device='cuda')
for staging in (128 << 20, 512 << 20, 2048 << 20):
    runner = sharded.ShardedCircuit(None, nl + 1, world, rank, bitops=[], run_stage=lambda st, sh: None,
                                    staging_bytes=staging)
    for _ in range(2):
        runner._exchange(shard, [0])
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        runner._exchange(shard, [0])
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        sent = (16 << nl) / 2
        print('EXCHANGE nl=%d staging=%d MiB  %.1f ms  %.0f GB/s per direction  env=%s' % (
            nl, staging >> 20, dt * 1e3, sent / dt / 1e9,
            {k: v for k, v in os.environ.items() if k.startswith('NCCL_')}), flush=True)
dist.destroy_process_group()

"""Sharded execution on real GPUs (NCCL over NVLink): needs >= 2 devices, skipped otherwise. The sharded result
is compared with the single-GPU engine and with the C oracle (the reference has no distributed path,
so parity is against the unsharded computation of the same circuit)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import AMP_TOL

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, depth, seed, out_dir, jit=False):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    if jit:
        os.environ['QFB_JIT'] = '1'      # sweep-specialised kernels: the remaps are pipelined slice by slice
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import quantumflow_b200 as qf
        from quantumflow_b200 import engine, sharded, workloads
        circ = workloads.wb_circuit(qf, n, depth, seed)
        # 256 KiB staging chunks: the in-place exchange runs its chunked, double-buffered path
        runner = sharded.ShardedCircuit(circ, n, world, rank, staging_bytes=1 << 18)
        nl = runner.nl
        shard = torch.zeros(1 << nl, dtype=torch.complex128, device='cuda')
        if rank == 0:
            shard[0] = 1
        shard = runner.execute(shard)
        if jit:
            # a second execution from the same start: the peer-memory barrier's epochs carry on
            again = torch.zeros_like(shard)
            if rank == 0:
                again[0] = 1
            again = runner.execute(again)
            assert torch.equal(again, shard)
        n2 = engine.norm2(shard)
        dist.all_reduce(n2)
        np.save(os.path.join(out_dir, 'shard{}.npy'.format(rank)), shard.cpu().numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, 'meta.npy'),
                    np.asarray(list(runner.final_phys_of) + [runner._remaps, runner._pipelined]))
            np.save(os.path.join(out_dir, 'norm.npy'), np.asarray([float(n2)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,depth,seed,jit', [(2, 20, 6, 0, False), (2, 22, 4, 1, False), (4, 21, 5, 2, False),
                                                     (8, 22, 4, 3, False), (2, 22, 8, 4, True), (4, 23, 6, 5, True),
                                                     (8, 24, 5, 6, True)])
def test_sharded_matches_single_gpu_and_oracle(tmp_path, world, n, depth, seed, jit):
    if torch.cuda.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    import torch.multiprocessing as mp
    from oracle import c_oracle
    from oracle import qf_oracle as O
    from quantumflow_b200 import sharded, workloads
    mp.spawn(_worker, args=(world, _free_port(), n, depth, seed, str(tmp_path), jit), nprocs=world, join=True)
    p = world.bit_length() - 1
    shards = [np.load(os.path.join(str(tmp_path), 'shard{}.npy'.format(r))) for r in range(world)]
    meta = list(np.load(os.path.join(str(tmp_path), 'meta.npy')))
    got = sharded.gather_logical(shards, n, p, [int(v) for v in meta[:n]])
    want = c_oracle.run_specs(workloads.wb_gate_list(n, depth, seed), n, O.gate_matrix)
    assert np.abs(got - want).max() < AMP_TOL
    assert abs(float(np.load(os.path.join(str(tmp_path), 'norm.npy'))[0]) - 1) < 1e-10
    assert int(meta[n]) >= 1
    if jit:
        assert int(meta[n + 1]) >= 1, 'no remap was pipelined'


def _readout_worker(rank, world, port, n, phys_of, phys, diag, uniforms, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from quantumflow_b200 import engine, sharded
        nl = n - (world.bit_length() - 1)
        shard = torch.from_numpy(np.ascontiguousarray(phys[rank << nl:(rank + 1) << nl])).cuda()
        before = engine.launch_count()
        vals = [sharded.norm2(shard)]
        for b in range(n):
            vals.extend(sharded.marginal(shard, b, phys_of, nl, rank))
        local_diag = torch.from_numpy(sharded.physical_diagonal(diag, phys_of, nl, rank)).cuda()
        vals.append(sharded.expectation_diag(shard, local_diag))
        samples = sharded.sample_indices(shard, uniforms, phys_of, nl, rank, world)
        assert engine.launch_count() > before
        np.save(os.path.join(out_dir, 'readout{}.npy'.format(rank)), np.concatenate([vals, samples]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,seed', [(1, 12, 0), (2, 13, 1), (4, 14, 2)])
def test_sharded_readout_on_device(tmp_path, world, n, seed):
    """SURVEY 8e 'Reductions / readout' with the device kernels as the per-rank reductions and NCCL for the few
    doubles that are combined (the CPU twin with numpy shards and gloo: tests/test_sharded_cpu.py)."""
    if torch.cuda.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    import torch.multiprocessing as mp
    from quantumflow_b200 import sharded
    rng = np.random.RandomState(seed)
    full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    full *= 0.7 / np.linalg.norm(full)
    phys_of = [int(v) for v in rng.permutation(n)]
    phys = sharded.scatter_physical(full, n, phys_of)
    diag = rng.normal(size=1 << n)
    uniforms = rng.random_sample(64)
    mp.spawn(_readout_worker, args=(world, _free_port(), n, phys_of, phys, diag, uniforms, str(tmp_path)),
             nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), 'readout{}.npy'.format(r))) for r in range(world)]
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    got = outs[0]
    probs = np.abs(full) ** 2
    assert abs(got[0] - probs.sum()) < 1e-13
    for b in range(n):
        one = ((np.arange(1 << n) >> b) & 1).astype(bool)
        assert abs(got[1 + 2 * b] - probs[~one].sum()) < 1e-13 and abs(got[2 + 2 * b] - probs[one].sum()) < 1e-13
    assert abs(got[1 + 2 * n] - (probs * diag).sum()) < 1e-12
    samples = got[2 + 2 * n:].astype(np.int64)
    cdf = np.cumsum(np.abs(phys) ** 2)
    physical = np.minimum(np.searchsorted(cdf, uniforms * cdf[-1], side='right'), cdf.size - 1)
    logical = np.zeros_like(physical)
    for b, pos in enumerate(phys_of):
        logical |= ((physical >> pos) & 1) << b
    assert (samples == logical).mean() > 0.9
    assert np.all(probs[samples] > 0)

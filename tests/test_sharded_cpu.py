"""Multi-GPU path, host side: the remap scheduler and the pairwise block exchange, run with torch.distributed
(gloo, world_size 2 and 4) on the CPU. The local kernels are replaced by the oracle (a test double handed to
ShardedCircuit), so what is under test is exactly what the reference does not have: which qubits become global,
the logical->physical bit map, and the exchange pattern."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qf_oracle as O
from quantumflow_b200 import planner, sharded, workloads

from conftest import AMP_TOL


def _bitops(specs, n):
    return [(O.gate_matrix(name, params), [n - 1 - q for q in qubits]) for name, params, qubits in specs]


def test_schedule_keeps_every_mixing_operator_local_and_tracks_the_map():
    n, p = 10, 2
    specs = workloads.wb_gate_list(n, 6, 3)
    ops = _bitops(specs, n)
    steps, phys_of = sharded.schedule(n, p, ops)
    assert sorted(phys_of) == list(range(n))
    nl = n - p
    nstage = 0
    gates = set()
    for st in steps:
        if isinstance(st, sharded.Stage):
            nstage += 1
            assert st.final_perm is None or sorted(st.final_perm) == list(range(nl))
            assert len(st.bitops) == len(st.items)
            for item, (mat, bits) in zip(st.items, st.bitops):
                mix, _ = sharded._mixing_and_diag_bits(np.asarray(mat), list(bits))
                assert all(b < nl for b in mix), 'mixing operator on a rank bit'
                gates.add(item.gate_index)          # a gate may contribute several phase terms
        else:
            assert 1 <= len(st.rank_positions) <= p
    assert gates == set(range(len(specs)))
    nremap = len(steps) - nstage
    assert 1 <= nremap <= 2 * 6 + 2          # about one remap per layer (SURVEY 8e estimate)
    # p = 0 degenerates to a single stage
    steps0, phys0 = sharded.schedule(n, 0, ops)
    assert len(steps0) == 1 and phys0 == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, specs, full_in, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        p = world.bit_length() - 1
        nl = n - p

        def run_stage(stage, shard):
            vec = shard.numpy()
            for mat, bits in stage.bitops:
                # rank bits only ever appear as diagonal / control bits: evaluate them from the rank
                k = len(bits)
                m = np.asarray(mat, dtype=np.complex128).reshape(1 << k, 1 << k)
                hi = [q for q in range(k) if bits[q] >= nl]
                if hi:
                    keep = [q for q in range(k) if bits[q] < nl]
                    sel = [slice(None)] * (2 * k)
                    t = m.reshape([2] * (2 * k))
                    for q in hi:
                        v = (rank >> (bits[q] - nl)) & 1
                        sel[q] = v
                        sel[k + q] = v
                    m = t[tuple(sel)].reshape(1 << len(keep), 1 << len(keep))
                    bits = [bits[q] for q in keep]
                    if not bits:
                        vec *= m[0, 0]
                        continue
                vec[:] = O.tensormul_flat(m, vec, list(bits))

        def permute(shard, perm, out):
            idx = np.arange(1 << nl, dtype=np.int64)
            src = np.zeros_like(idx)
            for j, pj in enumerate(perm):
                src |= ((idx >> j) & 1) << pj
            out.numpy()[:] = shard.numpy()[src]

        # a tiny staging buffer forces the chunked, double-buffered path of the in-place exchange
        runner = sharded.ShardedCircuit(None, n, world, rank, bitops=_bitops(specs, n), run_stage=run_stage,
                                        permute=permute, staging_bytes=16 * 24)
        shard = torch.from_numpy(np.ascontiguousarray(full_in[rank << nl:(rank + 1) << nl]))
        shard = runner.execute(shard)
        np.save(os.path.join(out_dir, 'shard{}.npy'.format(rank)), shard.numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, 'phys_of.npy'), np.asarray(runner.final_phys_of))
            np.save(os.path.join(out_dir, 'remaps.npy'), np.asarray([runner._remaps]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,depth,seed', [(2, 9, 5, 0), (4, 10, 4, 1), (2, 8, 3, 2)])
def test_sharded_execution_matches_oracle(tmp_path, world, n, depth, seed):
    specs = workloads.wb_gate_list(n, depth, seed)
    rng = np.random.RandomState(seed)
    full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    full /= np.linalg.norm(full)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, specs, full, str(tmp_path)), nprocs=world, join=True)
    p = world.bit_length() - 1
    shards = [np.load(os.path.join(str(tmp_path), 'shard{}.npy'.format(r))) for r in range(world)]
    phys_of = list(np.load(os.path.join(str(tmp_path), 'phys_of.npy')))
    got = sharded.gather_logical(shards, n, p, phys_of)
    want = O.run_specs(specs, n, full.reshape([2] * n)).reshape(-1)
    assert np.abs(got - want).max() < AMP_TOL
    assert int(np.load(os.path.join(str(tmp_path), 'remaps.npy'))[0]) >= 1


def _swap_index_bits(vec, pairs):
    idx = np.arange(vec.size, dtype=np.int64)
    src = idx.copy()
    for a, b in pairs:
        diff = ((idx >> a) ^ (idx >> b)) & 1
        src ^= (diff << a) | (diff << b)
    return vec[src]


@pytest.mark.parametrize('n,p,tile,depth,seed', [(12, 2, 7, 6, 0), (11, 1, 8, 8, 1), (13, 3, 8, 5, 2), (12, 2, 9, 7, 3)])
def test_sharded_plans_run_on_the_plan_emulator(n, p, tile, depth, seed):
    """The GPU path of a sharded circuit, minus the GPU: every stage is planned exactly as ShardedCircuit plans it
    (the scheduler's own sweep split, preset=, with the remap's local permutation fused into the last sweep's
    store or appended as a bare sweep), the plan blobs are validated by the library and executed by the plan
    emulator on every rank's shard with index_hi = rank, and a remap exchanges index bits of the concatenated
    shards as ShardedCircuit._exchange does."""
    import plan_emulator as E
    from test_planner import validate_with_library
    specs = workloads.wb_gate_list(n, depth, seed)
    nl = n - p
    steps, phys_of = sharded.schedule(n, p, _bitops(specs, n), tile_bits=tile)
    rng = np.random.RandomState(seed)
    full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    full /= np.linalg.norm(full)
    phys = full.copy()                                  # identity map at the start: physical = logical
    fused = bare = 0
    for st in steps:
        if isinstance(st, sharded.Stage):
            assert sum(count for count, _ in st.parts) == len(st.items)
            segs = planner.build_segments_from_items(nl, st.items, tile_bits=tile, final_perm=st.final_perm,
                                                     preset=st.parts)
            nsweeps = sum(s.nsweeps for s in segs)
            if st.final_perm is not None:
                fused += nsweeps == len(st.parts)
                bare += nsweeps > len(st.parts)
            for rank in range(1 << p):
                shard = phys[rank << nl:(rank + 1) << nl].copy()
                for seg in segs:
                    assert seg.kind == 'plan'
                    validate_with_library(seg.blob)
                    shard = E.execute(seg.blob, shard, rank)
                phys[rank << nl:(rank + 1) << nl] = shard
        else:
            k = len(st.rank_positions)
            phys = _swap_index_bits(phys, [(nl + t, nl - k + i) for i, t in enumerate(st.rank_positions)])
    shards = [phys[r << nl:(r + 1) << nl] for r in range(1 << p)]
    got = sharded.gather_logical(shards, n, p, phys_of)
    want = O.run_specs(specs, n, full.reshape([2] * n)).reshape(-1)
    assert np.abs(got - want).max() < AMP_TOL
    print('local permutations: fused into a stage sweep', fused, '/ bare sweep', bare)


@pytest.mark.parametrize('n,p', [(32, 2), (34, 1), (35, 2), (36, 3)])
def test_benchmark_size_sharded_plans_pass_the_library_validation(n, p):
    """The plans bench.py launches at 2 / 4 / 8 GPUs (33 qubits per GPU: 34 / 35 / 36 qubits, BASELINE.json
    configs[4]; and the 30-per-GPU shape of round 1), built as ShardedCircuit builds them, checked by the library's
    plan validation (the host half of qfb_plan_upload) and emitted + compiled for sm_100a as the sweep-specialised
    kernels (qfb_jit_check); too large to emulate."""
    import ctypes
    from quantumflow_b200 import _lib
    from test_planner import validate_with_library
    lib = _lib.load()
    specs = workloads.wb_gate_list(n, 20, 0)
    steps, phys_of = sharded.schedule(n, p, _bitops(specs, n))
    assert sorted(phys_of) == list(range(n))
    nsweeps = nremaps = 0
    for st in steps:
        if isinstance(st, sharded.Stage):
            segs = planner.build_segments_from_items(n - p, st.items, final_perm=st.final_perm, preset=st.parts)
            for seg in segs:
                assert seg.kind == 'plan'
                validate_with_library(seg.blob)
                log = ctypes.create_string_buffer(4096)
                assert lib.qfb_jit_check(seg.blob, len(seg.blob), log, len(log)) == 0, log.value.decode()
            nsweeps += sum(s.nsweeps for s in segs)
        else:
            nremaps += 1
    assert nsweeps <= 26 and 1 <= nremaps <= 6, (nsweeps, nremaps)


def _readout_worker(rank, world, port, n, phys_of, phys, diag, uniforms, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        p = world.bit_length() - 1
        nl = n - p
        shard = torch.from_numpy(np.ascontiguousarray(phys[rank << nl:(rank + 1) << nl]))

        def np_norm2(t):
            return float(np.vdot(t.numpy(), t.numpy()).real)

        def np_marginal(t, pos):
            pr = np.abs(t.numpy()) ** 2
            one = ((np.arange(pr.size) >> pos) & 1).astype(bool)
            return pr[~one].sum(), pr[one].sum()

        def np_expect(t, d):
            return float((np.abs(t.numpy()) ** 2 * d.numpy()).sum())

        def np_search(t, inner):
            cdf = np.cumsum(np.abs(t.numpy()) ** 2)
            return np.minimum(np.searchsorted(cdf, np.asarray(inner) * cdf[-1], side='right'), cdf.size - 1)

        res = {'norm2': sharded.norm2(shard, local=np_norm2)}
        res['marginals'] = [sharded.marginal(shard, b, phys_of, nl, rank, local=np_marginal, local_norm2=np_norm2)
                            for b in range(n)]
        local_diag = torch.from_numpy(sharded.physical_diagonal(diag, phys_of, nl, rank))
        res['expect'] = sharded.expectation_diag(shard, local_diag, local=np_expect)
        res['samples'] = sharded.sample_indices(shard, uniforms, phys_of, nl, rank, world, local_norm2=np_norm2,
                                                local_search=np_search)
        np.save(os.path.join(out_dir, 'readout{}.npy'.format(rank)),
                np.concatenate([[res['norm2']], np.ravel(res['marginals']), [res['expect']], res['samples']]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,seed', [(2, 8, 0), (4, 9, 1)])
def test_sharded_readout_reductions(tmp_path, world, n, seed):
    """SURVEY 8e 'Reductions / readout': norm, marginals (local and rank bits), a diagonal expectation and CDF
    sampling of a sharded state agree with the unsharded numpy computation, identically on every rank."""
    rng = np.random.RandomState(seed)
    p = world.bit_length() - 1
    full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    full *= 0.7 / np.linalg.norm(full)                  # unnormalised on purpose
    phys_of = [int(v) for v in rng.permutation(n)]      # some logical bits live in the rank bits
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
    assert abs(got[0] - probs.sum()) < 1e-14
    for b in range(n):
        one = ((np.arange(1 << n) >> b) & 1).astype(bool)
        assert abs(got[1 + 2 * b] - probs[~one].sum()) < 1e-14 and abs(got[2 + 2 * b] - probs[one].sum()) < 1e-14
    assert abs(got[1 + 2 * n] - (probs * diag).sum()) < 1e-13
    samples = got[2 + 2 * n:].astype(np.int64)
    # the sampler inverts the CDF in PHYSICAL order: same construction on the gathered physical vector
    cdf = np.cumsum(np.abs(phys) ** 2)
    physical = np.minimum(np.searchsorted(cdf, uniforms * cdf[-1], side='right'), cdf.size - 1)
    logical = np.zeros_like(physical)
    for b, pos in enumerate(phys_of):
        logical |= ((physical >> pos) & 1) << b
    agree = samples == logical
    assert agree.mean() > 0.9                           # rounding at span edges may move a draw to a neighbour
    assert np.all(probs[samples] > 0)


def test_thin_sweep_with_nothing_to_exchange_is_taken_instead_of_raising():
    """Reproducer of a scheduler crash (round-1 review): many 2-qubit gates on the low bits, one layer over every
    other local bit, then H on the global bit. The stage used to end before the thin sweep while Belady kept the
    same global bit, so the remap exchanged nothing ('scheduler made no progress')."""
    import random
    rnd = random.Random(5)
    for n in (16, 17):
        p, nl = 1, n - 1
        ops = []
        for _ in range(200):
            a, b = rnd.sample(range(12), 2)
            ops.append((O.gate_matrix('CAN', (0.1, 0.2, 0.3)), [a, b]))
        free = list(range(3, nl))
        for a, b in zip(free[::2], free[1::2]):
            ops.append((O.gate_matrix('CAN', (0.3, 0.1, 0.2)), [a, b]))
        ops.append((O.gate_matrix('H', ()), [n - 1]))
        for thin in (0.0, 0.6, 0.9):
            steps, phys_of, _ = sharded._schedule_once(n, p, ops, None, None, None, thin)
            assert sorted(phys_of) == list(range(n))
            done = sum(len(st.items) for st in steps if isinstance(st, sharded.Stage))
            assert done >= len(ops)
            for st in steps:
                if isinstance(st, sharded.Stage):
                    for mat, bits in st.bitops:
                        mix, _ = sharded._mixing_and_diag_bits(np.asarray(mat), list(bits))
                        assert all(b < nl for b in mix)
        steps, phys_of = sharded.schedule(n, p, ops)
        assert any(isinstance(st, sharded.Remap) for st in steps)


# ---------------------------------------------------------------------------------------------------------
# pipelined remaps: the host logic that chooses slices and chains (the kernels are covered by test_gpu_sharded.py)
# ---------------------------------------------------------------------------------------------------------

class _FakePlan:
    """Stands in for engine.UploadedPlan: sweep count, non-tile masks, and a record of what was launched."""

    def __init__(self, nontile_masks, log, name):
        self._masks, self._log, self._name = list(nontile_masks), log, name
        self.specialised = True

    @property
    def nsweeps(self):
        return len(self._masks)

    def nontile_mask(self, sweep):
        return self._masks[sweep]

    def launch_part(self, tensor, first, count, index_hi=0, fix_mask=0, fix_value=0, ctas_per_sm=0):
        self._log.append((self._name, first, count, fix_mask, fix_value))


def _fake_runner(nl, p, masks_per_stage, log):
    """A ShardedCircuit shell whose stages are one fake plan each (no planning, no GPU)."""
    runner = sharded.ShardedCircuit.__new__(sharded.ShardedCircuit)
    runner.nl, runner.p, runner.world, runner.rank = nl, p, 1 << p, 0
    stages = []
    for i, masks in enumerate(masks_per_stage):
        st = sharded.Stage([], None, None)
        seg = planner.Segment('plan', blob=b'', nsweeps=len(masks))
        seg.uploaded = _FakePlan(masks, log, 'stage%d' % i)
        st.segments = [seg]
        stages.append(st)
    return runner, stages


def test_pipeline_shape_takes_common_non_tile_bits_below_the_exchanged_blocks(monkeypatch):
    monkeypatch.delenv('QFB_REMAP_SLICE_BITS', raising=False)
    monkeypatch.delenv('QFB_REMAP_CHAIN', raising=False)
    nl, p = 30, 1
    bit = lambda *pos: sum(1 << b for b in pos)                                   # noqa: E731
    log = []
    # stage 0: 4 sweeps; stage 1: 5 sweeps. Bits 20..24 are outside every tile, 28 only outside the nearest sweeps,
    # 29 is the exchanged block bit region (nl - k - 1 = 28 is the half-block split: not allowed either)
    s0 = [bit(20, 21, 22, 23, 24), bit(20, 21, 22, 23, 24, 27), bit(20, 21, 22, 23, 24, 27), bit(20, 21, 22, 23, 24, 27, 28)]
    s1 = [bit(20, 21, 22, 23, 24, 27, 28), bit(20, 21, 22, 23, 24, 27), bit(5, 20, 21, 22, 23, 24), bit(20, 21), bit(20, 21)]
    runner, (a, b) = _fake_runner(nl, p, [s0, s1], log)
    remap = sharded.Remap([0])
    bits, da, db = runner._pipeline_shape(a, remap, b, 0, False)
    assert (da, db) == (3, 3)                       # three sweeps per side keep 3 common selector bits
    assert bits == [22, 23, 24] and all(12 <= x < nl - 1 - 1 for x in bits)
    # the next stage feeds another remap: it keeps at least half of its sweeps for that one
    bits2, da2, db2 = runner._pipeline_shape(a, remap, b, 0, True)
    assert db2 == 2 and da2 == 3
    # sweeps that the previous pipeline has already run are not available
    assert runner._pipeline_shape(a, remap, b, 2, False)[1] == 2
    assert runner._pipeline_shape(a, remap, b, 4, False) is None
    # one selector bit (2 slices) only when nothing better exists; below bit 12 never
    only = [bit(3, 4, 13)]
    runner2, (c, d) = _fake_runner(nl, p, [only, only], [])
    assert runner2._pipeline_shape(c, remap, d, 0, False) == ([13], 1, 1)
    runner3, (e, f) = _fake_runner(nl, p, [[bit(3, 4)], [bit(3, 4)]], [])
    assert runner3._pipeline_shape(e, remap, f, 0, False) is None
    monkeypatch.setenv('QFB_REMAP_CHAIN', '1')
    assert runner._pipeline_shape(a, remap, b, 0, False)[1:] == (1, 1)


def test_stage_parts_skip_exactly_the_sweeps_the_pipelines_run():
    log = []
    runner, (a,) = _fake_runner(30, 1, [[1, 1, 1, 1, 1]], log)
    runner._run_stage_part(a, None, 2, 2)
    assert log == [('stage0', 2, 1, 0, 0)]
    log.clear()
    runner._run_stage_part(a, None, 0, 0)
    assert log == [('stage0', 0, 5, 0, 0)]
    log.clear()
    runner._run_stage_part(a, None, 3, 2)          # nothing left between the two pipelines
    assert log == []


@pytest.mark.parametrize('n,p', [(34, 1), (35, 2), (36, 3)])
def test_overlapped_execution_launches_every_sweep_of_the_benchmark_schedules_exactly_once(n, p, monkeypatch, tmp_path):
    """Dry run of ShardedCircuit._execute_overlapped on the schedules bench.py runs at 2 / 4 / 8 GPUs (33 qubits per
    GPU), without a GPU: the real scheduler, stage plans and pipeline shapes (_pipeline_shape on the plans' own
    non-tile masks), with recording stand-ins for the uploaded plans and for the device half of a pipelined remap.
    Every sweep of every stage must be launched exactly once over the whole shard -- in one piece by _run_stage_part
    or slice by slice (all 2^v values of the selector bits) inside a remap's pipeline -- in stage order, and every
    remap must happen once."""
    import plan_emulator as E
    monkeypatch.delenv('QFB_REMAP_SLICE_BITS', raising=False)
    monkeypatch.delenv('QFB_REMAP_CHAIN', raising=False)
    specs = workloads.wb_gate_list(n, 20, 0)
    runner = sharded.ShardedCircuit(None, n, 1 << p, 0, bitops=_bitops(specs, n))
    log = []
    stage_names = {}
    for si, st in enumerate(s for s in runner.steps if isinstance(s, sharded.Stage)):
        for gi, seg in enumerate(st.segments):
            assert seg.kind == 'plan'
            parsed = E.parse(seg.blob)
            masks = [sum(1 << b for b in sw['hole']) for sw in parsed['sweeps']]
            assert len(masks) == seg.nsweeps
            name = 'stage{}.{}'.format(si, gi)
            seg.uploaded = _FakePlan(masks, log, name)
            stage_names[name] = seg.nsweeps
    remaps = []
    monkeypatch.setattr(runner, '_map_peers', lambda shard: None)
    monkeypatch.setattr(runner, '_exchange_peer', lambda shard, pos: remaps.append(('whole', tuple(pos))))

    import shutil
    import subprocess
    from test_jit_emulated import sweep_source
    ptxas = shutil.which('ptxas') or ('/usr/local/cuda/bin/ptxas' if os.path.exists('/usr/local/cuda/bin/ptxas') else None)
    variants = []

    def fake_pipeline(shard, prev, remap, nxt, bits, da, db):
        a, b = prev.segments[-1].uploaded, nxt.segments[0].uploaded
        # the slice variants this pipeline launches (compiled on first use on the device): generate them here and, where
        # the toolkit's assembler is at hand, assemble them for sm_100a
        fix = sum(1 << pos for pos in bits)
        for seg, sweeps in ((prev.segments[-1], range(a.nsweeps - da, a.nsweeps)), (nxt.segments[0], range(db))):
            for i in sweeps:
                ptx = sweep_source(seg.blob, i, fix)[0]
                assert 'p_fix' in ptx
                variants.append(ptx)
        assert all((a.nontile_mask(i) >> pos) & 1 for i in range(a.nsweeps - da, a.nsweeps) for pos in bits)
        assert all((b.nontile_mask(i) >> pos) & 1 for i in range(db) for pos in bits)
        k = len(remap.rank_positions)
        assert all(12 <= pos < runner.nl - k - 1 for pos in bits) and bits == sorted(bits) and 1 <= len(bits) <= 3
        mask = sum(1 << pos for pos in bits)
        values = [sum(((sl >> t) & 1) << pos for t, pos in enumerate(bits)) for sl in range(1 << len(bits))]
        for v in values:
            a.launch_part(shard, a.nsweeps - da, da, fix_mask=mask, fix_value=v)
        remaps.append(('pipelined', tuple(remap.rank_positions)))
        for v in values:
            b.launch_part(shard, 0, db, fix_mask=mask, fix_value=v)

    monkeypatch.setattr(runner, '_remap_pipelined', fake_pipeline)
    runner._execute_overlapped(None)
    # coverage per sweep: 1.0 for a whole launch, 2^-v per slice launch
    covered = {name: [0.0] * count for name, count in stage_names.items()}
    order = []
    for name, first, count, fix_mask, _value in log:
        share = 1.0 / (1 << bin(fix_mask).count('1'))
        for i in range(first, first + count):
            covered[name][i] += share
        order.append(name)
    for name, shares in covered.items():
        assert all(abs(s - 1.0) < 1e-12 for s in shares), (name, shares)
    # stages run in order (a pipeline interleaves two neighbouring stages only)
    stage_of = [int(name[5:].split('.')[0]) for name in order]
    assert all(b >= a - 1 for a, b in zip(stage_of, stage_of[1:]))
    assert len(remaps) == sum(isinstance(s, sharded.Remap) for s in runner.steps)
    assert runner._remaps == len(remaps) and runner._executions == 1
    assert len(variants) == runner._pipelined_sweeps
    if ptxas is not None:
        for k, ptx in enumerate(variants):
            path = os.path.join(str(tmp_path), 'variant{}.ptx'.format(k))
            with open(path, 'w') as f:
                f.write(ptx)
            done = subprocess.run([ptxas, '-arch=sm_100a', '-O3', path, '-o', path + '.cubin'], capture_output=True,
                                  text=True)
            assert done.returncode == 0, done.stderr
    print(n, p, 'remaps', remaps, 'sweeps inside pipelines', runner._pipelined_sweeps, 'of', sum(stage_names.values()),
          '| slice variants assembled:', len(variants) if ptxas else 'no ptxas')


class _FakeShard:
    """What ShardedCircuit._swap_runs reads of a shard tensor: size and a 'device address'."""

    def __init__(self, base, numel):
        self._base, self._numel = base, numel

    def numel(self):
        return self._numel

    def data_ptr(self):
        return self._base

    def element_size(self):
        return 16


@pytest.mark.parametrize('p,nl,rank_positions', [(1, 6, [0]), (2, 7, [0, 1]), (2, 6, [1]), (3, 7, [0, 2]), (3, 8, [0, 1, 2])])
def test_peer_exchange_runs_swap_exactly_the_remapped_blocks(p, nl, rank_positions):
    """The address lists of the peer-memory exchange (ShardedCircuit._swap_runs -> qfb_remap_swap: of every pair's
    block the lower rank swaps the first half, the higher rank the second half), executed on numpy shards for every
    rank, equal the index-bit swap a remap stands for; the slice form (qfb_remap_swap_slice: only offsets of a run
    whose selector bits equal the slice value) covers the same amplitudes once over all slice values."""
    world = 1 << p
    k = len(rank_positions)
    rng = np.random.RandomState(p * 100 + nl)
    full = rng.normal(size=1 << (nl + p)) + 1j * rng.normal(size=1 << (nl + p))
    want = _swap_index_bits(full, [(nl + t, nl - k + i) for i, t in enumerate(rank_positions)])
    stride = 1 << 40                                       # rank r's shard "lives" at r * stride

    def runs_of(rank):
        runner = sharded.ShardedCircuit.__new__(sharded.ShardedCircuit)
        runner.rank, runner.world, runner._comm_bytes = rank, world, 0
        runner._peers = [r * stride for r in range(world)]
        return runner._swap_runs(_FakeShard(rank * stride, 1 << nl), rank_positions)

    def swap(shards, local, remote, lo, hi):
        (ra, oa), (rb, ob) = divmod(local, stride), divmod(remote, stride)
        a = shards[ra][oa // 16 + lo: oa // 16 + hi].copy()
        shards[ra][oa // 16 + lo: oa // 16 + hi] = shards[rb][ob // 16 + lo: ob // 16 + hi]
        shards[rb][ob // 16 + lo: ob // 16 + hi] = a

    whole = [full[r << nl:(r + 1) << nl].copy() for r in range(world)]
    sliced = [s.copy() for s in whole]
    selpos = [2, 3] if nl - k - 1 > 3 else [1]               # selector bits below the half-block split
    total = 0
    for rank in range(world):
        local, remote, counts = runs_of(rank)
        assert len(local) == (1 << k) - 1
        for la, ra, n in zip(local, remote, counts):
            assert la // stride == rank and ra // stride != rank
            swap(whole, la, ra, 0, n)
            total += n
            for value in range(1 << len(selpos)):
                for off in range(n):
                    if all(((off >> pos) & 1) == ((value >> t) & 1) for t, pos in enumerate(selpos)):
                        swap(sliced, la, ra, off, off + 1)
    assert total == world * ((1 << nl) - (1 << (nl - k))) // 2        # every off-diagonal block once, half per rank
    assert np.array_equal(np.concatenate(whole), want)
    assert np.array_equal(np.concatenate(sliced), want)

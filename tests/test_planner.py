"""Planner + binary plan format, verified on the CPU with the plan emulator (tests/plan_emulator.py walks a
plan the way sweep_kernel does) against the oracle. The GPU tests then only have to establish that the kernel
implements the same walk."""
import math
import random

import numpy as np
import pytest

import plan_emulator as E
from oracle import qf_oracle as O
from quantumflow_b200 import planner, workloads

from conftest import AMP_TOL


def bitops_of(specs, n):
    return [(O.gate_matrix(name, params), [n - 1 - q for q in qubits]) for name, params, qubits in specs]


def validate_with_library(blob):
    """The library's own structural validation (host code of qfb_plan_upload; no GPU needed)."""
    import ctypes
    from quantumflow_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib.qfb_plan_validate(ctypes.cast(buf, ctypes.c_void_p), len(blob))
    assert rc == 0, lib.qfb_last_error()


def run_segments(segments, state, index_hi=0):
    state = np.array(state, dtype=np.complex128).reshape(-1)
    for seg in segments:
        if seg.kind == 'plan':
            validate_with_library(seg.blob)
            state = E.execute(seg.blob, state, index_hi)
        else:
            state = O.tensormul_flat(seg.mat, state, list(seg.bits))
    return state


def zero(n):
    v = np.zeros(1 << n, dtype=np.complex128)
    v[0] = 1
    return v


@pytest.mark.parametrize('n,depth,seed,tile', [(12, 20, 0, 12), (12, 6, 1, 8), (11, 5, 2, 11), (10, 5, 2, 10),
                                                (9, 4, 3, 6), (8, 6, 4, 6), (13, 3, 5, 13), (7, 9, 6, 7)])
def test_wb_plan_matches_oracle(n, depth, seed, tile):
    specs = workloads.wb_gate_list(n, depth, seed)
    segments = planner.build_segments(n, bitops_of(specs, n), tile_bits=tile)
    assert all(s.kind == 'plan' for s in segments)
    got = run_segments(segments, zero(n))
    want = O.run_specs(specs, n).reshape(-1)
    assert np.abs(got - want).max() < AMP_TOL
    stats = planner.plan_stats(segments)
    assert stats['ops'] == len(specs) and stats['sweeps'] < len(specs) / 4


def test_plan_matches_reference_fixture(golden):
    want = golden('workloads.npz')['wb12_seed1']
    specs = workloads.wb_gate_list(12, 20, 1)
    got = run_segments(planner.build_segments(12, bitops_of(specs, 12)), zero(12))
    assert np.abs(got - want).max() < AMP_TOL


def test_controls_three_qubit_gates_and_fallback(golden):
    """CCNOT/CSWAP become controlled 1-/2-bit ops, CAN/PISWAP dense 2-bit ops, ISWAP/SWAP permutation ops."""
    rnd = random.Random(11)
    specs = []
    names1 = ['H', 'S', 'T', 'X', 'Y', 'Z', 'S_H', 'T_H']
    for d in range(12):
        for q in range(9):
            specs.append((rnd.choice(names1), (), (q,)))
        a, b, c = rnd.sample(range(9), 3)
        specs.append(('CCNOT', (), (a, b, c)))
        a, b, c = rnd.sample(range(9), 3)
        specs.append(('CSWAP', (), (a, b, c)))
        a, b = rnd.sample(range(9), 2)
        specs.append(('CAN', (rnd.random(), rnd.random(), rnd.random()), (a, b)))
        a, b = rnd.sample(range(9), 2)
        specs.append(('PISWAP', (rnd.random(),), (a, b)))
        a, b = rnd.sample(range(9), 2)
        specs.append(('ISWAP', (), (a, b)))
        a, b = rnd.sample(range(9), 2)
        specs.append(('CPHASE', (rnd.random(),), (a, b)))
        a, b = rnd.sample(range(9), 2)
        specs.append(('SWAP', (), (a, b)))
        specs.append(('TX', (rnd.random(),), (rnd.randrange(9),)))
        specs.append(('ZYZ', (rnd.random(), rnd.random(), rnd.random()), (rnd.randrange(9),)))
    want = golden('workloads.npz')['mixed9_seed11']          # produced by the reference with the same `random`
    for tile in (9, 7, 6):
        segments = planner.build_segments(9, bitops_of(specs, 9), tile_bits=tile)
        got = run_segments(segments, zero(9))
        assert np.abs(got - want).max() < AMP_TOL, tile
    # a dense 3-qubit operator cannot be expressed by the executor: it becomes its own segment
    rng = np.random.RandomState(3)
    u3 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
    ops = bitops_of(specs[:20], 9) + [(u3, [7, 2, 4])] + bitops_of(specs[20:40], 9)
    segments = planner.build_segments(9, ops)
    assert [s.kind for s in segments] == ['plan', 'op', 'plan']
    want = zero(9)
    for m, bits in ops:
        want = O.tensormul_flat(m, want, bits)
    assert np.abs(run_segments(segments, zero(9)) - want).max() < AMP_TOL


def test_density_bitops_plan(golden):
    """Circuit.evolve as bit-level ops on the 2N-bit vector: U on ket bits, conj(U) on bra bits, channels as
    2-bit superoperators (reference ops.py:354-363)."""
    n = 6
    specs = workloads.wd_gate_list(n, 20, 0)
    ops = []
    for name, params, qubits in specs:
        ket_bits = [2 * n - 1 - q for q in qubits]
        bra_bits = [n - 1 - q for q in qubits]
        if name == 'DEPOLARIZING':
            ops.append((O.depolarizing_superop(params[0]), ket_bits + bra_bits))
        else:
            u = O.gate_matrix(name, params)
            ops.append((u, ket_bits))
            ops.append((u.conj(), bra_bits))
    segments = planner.build_segments(2 * n, ops)
    got = run_segments(segments, zero(2 * n)).reshape(64, 64)
    want = golden('workloads.npz')['wd6_seed0_kraus']
    assert np.abs(got - want).max() < AMP_TOL
    assert planner.plan_stats(segments)['sweeps'] <= 12


def test_sharded_plan_uses_rank_bits_for_diagonals_and_controls():
    """With the top p bits held in the rank (index_hi), diagonal operators and controls on those bits need no
    communication: the plan resolves them from index_hi."""
    n, p = 11, 2
    nl = n - p
    rnd = random.Random(5)
    specs = []
    for q in range(n):
        specs.append(('H', (), (q,)))
    for _ in range(40):
        kind = rnd.choice(['T', 'RZ', 'CZ', 'CNOT', 'ZZ', 'CCNOT', 'RX'])
        if kind in ('T', 'RZ'):
            q = rnd.randrange(n)
            specs.append((kind, (rnd.random(),) if kind == 'RZ' else (), (q,)))
        elif kind == 'RX':
            specs.append((kind, (rnd.random(),), (rnd.randrange(p, n),)))          # local target only
        elif kind in ('CZ', 'ZZ'):
            a, b = rnd.sample(range(n), 2)
            specs.append((kind, (rnd.random(),) if kind == 'ZZ' else (), (a, b)))
        elif kind == 'CNOT':
            c = rnd.randrange(n)
            t = rnd.choice([q for q in range(p, n) if q != c])
            specs.append((kind, (), (c, t)))
        else:
            t = rnd.randrange(p, n)
            c0, c1 = rnd.sample([q for q in range(n) if q != t], 2)
            specs.append((kind, (), (c0, c1, t)))
    # the initial H layer on the global qubits cannot be done locally: start from a random full state instead
    specs = specs[n:]
    rng = np.random.RandomState(0)
    full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    want = O.run_specs(specs, n, full.reshape([2] * n)).reshape(-1)
    ops = bitops_of(specs, n)                     # bit positions in the FULL index; >= nl means rank bit
    segments = planner.build_segments(nl, ops)
    got = np.empty_like(full)
    for rank in range(1 << p):
        shard = full[rank << nl: (rank + 1) << nl]
        got[rank << nl: (rank + 1) << nl] = run_segments(segments, shard, index_hi=rank)
    assert np.abs(got - want).max() < AMP_TOL
    with pytest.raises(ValueError):
        planner.build_segments(nl, [(O.gate_matrix('H'), [n - 1])])      # mixing a rank bit needs a remap


def test_layout_invariants():
    specs = workloads.wb_gate_list(14, 6, 7)
    segments = planner.build_segments(14, bitops_of(specs, 14))
    plan = E.parse(segments[0].blob)
    assert plan['M'] == 12
    worst = 1
    for sweep in plan['sweeps']:
        assert sweep['gpos'][:3] == [0, 1, 2]                       # low bits always in the tile
        assert sweep['gpos'] == sorted(sweep['gpos'])
        nr = len(sweep['rounds'])
        for r, rd in enumerate(sweep['rounds']):
            if r in (0, nr - 1):
                assert min(rd['regpos']) >= 3                       # lanes keep the low bits on rounds touching HBM
                assert rd['thrpos'][:3] == [0, 1, 2]
            worst = max(worst, E.conflict_degree(rd, plan['M']))
    assert worst == 1                                               # swizzle + lane assignment: conflict free
    # executes correctly with 4 tiles
    got = run_segments(segments, zero(14))
    assert np.abs(got - O.run_specs(specs, 14, flat=True).reshape(-1)).max() < AMP_TOL


def test_identity_and_global_phase_ops():
    ops = [(np.eye(2), [0]), (np.exp(0.3j) * np.eye(2), [3]), (O.gate_matrix('H'), [2]), (np.eye(4), [1, 2])]
    segments = planner.build_segments(6, ops)
    got = run_segments(segments, zero(6))
    want = O.tensormul_flat(O.gate_matrix('H'), zero(6), [2]) * np.exp(0.3j)
    assert np.abs(got - want).max() < 1e-15
    assert planner.build_segments(6, [(np.eye(2), [0])]) == []


def test_planner_rejects_tiny_states():
    with pytest.raises(ValueError):
        planner.build_segments(3, [(O.gate_matrix('H'), [0])])


@pytest.mark.parametrize('nbits,tile,perm_seed', [(9, 7, 0), (10, 9, 1), (12, 12, 2), (8, 6, 3), (13, 12, 4)])
def test_final_bit_permutation_is_fused_or_appended(nbits, tile, perm_seed):
    """build_segments(final_perm=...) ends the plan with the in-place bit permutation of a qubit remap: fused into
    the last sweep when the moved bits are tile bits, appended as bare sweeps otherwise."""
    rnd = random.Random(perm_seed)
    specs = workloads.wb_gate_list(nbits, 2, perm_seed)
    ops = [(O.gate_matrix(name, params), [nbits - 1 - q for q in qubits]) for name, params, qubits in specs]
    perm = list(range(nbits))
    moved = rnd.sample(range(nbits), min(nbits, 2 + 2 * (perm_seed % 3)))     # includes low bits sometimes
    shuffled = list(moved)
    rnd.shuffle(shuffled)
    for a, b in zip(moved, shuffled):
        perm[a] = b
    rng = np.random.RandomState(perm_seed)
    vec = rng.normal(size=1 << nbits) + 1j * rng.normal(size=1 << nbits)
    for with_ops in (True, False):
        segs = planner.build_segments(nbits, ops if with_ops else [], tile_bits=tile, final_perm=perm)
        got = vec.copy()
        for seg in segs:
            assert seg.kind == 'plan'
            validate_with_library(seg.blob)
            got = E.execute(seg.blob, got)
        want = vec.copy()
        if with_ops:
            want = O.run_specs(specs, nbits, want.reshape([2] * nbits)).reshape(-1)
        idx = np.arange(1 << nbits)
        src = np.zeros_like(idx)
        for j, pj in enumerate(perm):
            src |= ((idx >> j) & 1) << pj
        want = want[src]
        assert np.abs(got - want).max() < AMP_TOL


def _controlled(mat, nctrl=1):
    k = int(np.log2(mat.shape[0]))
    full = np.eye(1 << (k + nctrl), dtype=np.complex128)
    full[-(1 << k):, -(1 << k):] = mat
    return full


@pytest.mark.parametrize('seed,tile', [(0, 9), (1, 7), (2, 6), (3, 9)])
def test_frame_pass_flips_scales_and_controls(seed, tile):
    """absorb_frame under stress: X / Y flips, rotations close to a half turn (large LDU scales, the scale limit),
    controlled rotations with flipped controls, real X-shaped 2-bit operators, dense 1-bit and 2-bit operators,
    diagonal gates - all interleaved on 9 bits, executed by the plan emulator, against the plain oracle."""
    rnd = random.Random(seed)
    rng = np.random.RandomState(seed)
    n = 9
    ops = []
    for _ in range(140):
        kind = rnd.randrange(12)
        a, b, c = rnd.sample(range(n), 3)
        if kind == 0:
            ops.append((O.gate_matrix('X'), [a]))
        elif kind == 1:
            ops.append((O.gate_matrix('Y'), [a]))
        elif kind == 2:
            ops.append((O.gate_matrix('H'), [a]))
        elif kind == 3:       # rotations near pi: tiny cos(theta / 2), huge pending scales
            ops.append((O.gate_matrix(rnd.choice(['RX', 'RY']), (math.pi + rnd.uniform(-1e-3, 1e-3),)), [a]))
        elif kind == 4:
            ops.append((O.gate_matrix(rnd.choice(['RX', 'RY', 'RZ']), (rnd.uniform(0, 2 * math.pi),)), [a]))
        elif kind == 5:
            ops.append((O.gate_matrix('T'), [a]))
        elif kind == 6:
            ops.append((O.gate_matrix(rnd.choice(['CNOT', 'CZ'])), [a, b]))
        elif kind == 7:       # controlled rotation: the frame pass rewrites flipped controls
            ops.append((_controlled(O.gate_matrix('RY', (rnd.uniform(0, 6),))), [a, b]))
        elif kind == 8:
            ops.append((O.gate_matrix('CCNOT'), [a, b, c]))
        elif kind == 9:       # real X-shaped 2-bit operator (a Pauli-channel superoperator looks like this)
            m = np.zeros((4, 4))
            m[0, 0], m[0, 3], m[3, 0], m[3, 3], m[1, 1], m[1, 2], m[2, 1], m[2, 2] = rng.normal(size=8)
            ops.append((m.astype(np.complex128), [a, b]))
        elif kind == 10:      # dense 1-bit and 2-bit operators
            q = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
            ops.append((q, [a]))
        else:
            q = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
            ops.append((q, [a, b]))
    vec = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    vec /= np.linalg.norm(vec)
    want = vec.copy()
    for m, bits in ops:
        want = O.tensormul_flat(m, want, bits)
    got = run_segments(planner.build_segments(n, ops, tile_bits=tile), vec)
    assert np.abs(got - want).max() < AMP_TOL * max(1.0, np.abs(want).max())


def test_classified_items_round_trip_to_operators():
    """planner.item_bitop (what reference executors of sharded stages apply) reproduces the classified operator."""
    n = 6
    specs = workloads.wb_gate_list(n, 4, 7) + [('CCNOT', (), (0, 3, 5)), ('CPHASE', (0.3,), (1, 4))]
    ops = bitops_of(specs, n)
    vec = np.random.RandomState(0).normal(size=1 << n) + 0j
    want = vec.copy()
    for m, bits in ops:
        want = O.tensormul_flat(m, want, bits)
    got = vec.copy()
    for item in planner.classify_all(ops):
        m, bits = planner.item_bitop(item)
        got = O.tensormul_flat(m, got, bits)
    assert np.abs(got - want).max() < AMP_TOL


@pytest.mark.parametrize('seed', range(6))
def test_native_tile_search_matches_its_python_statement(seed):
    """qfb_plan_count_executed / qfb_plan_refine_tile (csrc/qfb_planhost.cu) against tests/plan_emulator.py on
    operator lists of benchmark circuits and on random masks, with and without forbidden / kept bits and caps."""
    import ctypes
    from quantumflow_b200 import _lib
    lib = _lib.load()
    rnd = random.Random(seed)
    nbits = rnd.choice([10, 14, 20, 33])
    if seed % 2 == 0:
        specs = workloads.wb_gate_list(nbits, 6, seed)
        items = planner.classify_all(bitops_of(specs, nbits))
        recs = [(op.mixmask, op.diagmask, float(op.cost), int(op.plan_bytes)) for op in items]
    else:
        recs = []
        for _ in range(300):
            mm = sum(1 << b for b in rnd.sample(range(nbits), rnd.choice([0, 1, 1, 2])))
            dd = sum(1 << b for b in rnd.sample(range(nbits), rnd.choice([0, 0, 1, 2]))) & ~mm
            recs.append((mm, dd, rnd.choice([0.0, 0.25, 1.0]), rnd.choice([32, 160, 576])))
    n = len(recs)
    mix = np.array([r[0] for r in recs], dtype=np.uint64)
    diag = np.array([r[1] for r in recs], dtype=np.uint64)
    cost = np.array([r[2] for r in recs], dtype=np.float64)
    nbytes = np.array([r[3] for r in recs], dtype=np.uint32)
    for trial in range(8):
        low = 7
        fmask = sum(1 << b for b in rnd.sample(range(3, nbits), rnd.choice([0, 0, 2]))) if nbits > 6 else 0
        free = [b for b in range(3, nbits) if not (fmask >> b) & 1]
        tile_bits = rnd.sample(free, min(len(free), rnd.choice([3, 5, 9])))
        tmask = low | sum(1 << b for b in tile_bits)
        keep = low | (sum(1 << b for b in tile_bits[:2]) if trial % 3 == 0 else 0)
        max_cost = rnd.choice([6.0, 28.0, 1e9])
        room = rnd.choice([2000, 40000])
        got = ctypes.c_int(-1)
        assert lib.qfb_plan_count_executed(mix.ctypes.data, diag.ctypes.data, cost.ctypes.data, nbytes.ctypes.data, n,
                                           tmask, fmask, max_cost, room, ctypes.byref(got)) == 0
        assert got.value == E.count_executed(recs, tmask, fmask, max_cost, room)
        out, cnt = ctypes.c_uint64(0), ctypes.c_int(-1)
        assert lib.qfb_plan_refine_tile(mix.ctypes.data, diag.ctypes.data, cost.ctypes.data, nbytes.ctypes.data, n,
                                        nbits, tmask, fmask, keep, max_cost, room, 6, ctypes.byref(out),
                                        ctypes.byref(cnt)) == 0
        want_mask, want_cnt = E.refine_tile(recs, nbits, tmask, fmask, keep, max_cost, room, 6)
        assert (out.value, cnt.value) == (want_mask, want_cnt)
        assert bin(out.value).count('1') == bin(tmask).count('1') and out.value & keep == keep
        assert not out.value & fmask
        # two-sweep look-ahead: whatever it scores, the result is a tile of the same size that keeps `keep`, avoids
        # `fmask`, and does not score worse than its start; same answer when asked twice
        m = bin(tmask).count('1')
        la, sc0, sc1 = ctypes.c_uint64(0), ctypes.c_int(-1), ctypes.c_int(-1)
        args = (mix.ctypes.data, diag.ctypes.data, cost.ctypes.data, nbytes.ctypes.data, n, nbits, 3, m)
        assert lib.qfb_plan_refine_tile_lookahead(*args, tmask, fmask, keep, max_cost, room, 6, 0, ctypes.byref(la),
                                                  ctypes.byref(sc0)) == 0 and la.value == tmask
        assert lib.qfb_plan_refine_tile_lookahead(*args, tmask, fmask, keep, max_cost, room, 6, 3, ctypes.byref(la),
                                                  ctypes.byref(sc1)) == 0
        assert bin(la.value).count('1') == m and la.value & keep == keep and not la.value & fmask
        assert sc1.value >= sc0.value >= E.count_executed(recs, tmask, fmask, max_cost, room)
        again = ctypes.c_uint64(0)
        assert lib.qfb_plan_refine_tile_lookahead(*args, tmask, fmask, keep, max_cost, room, 6, 3,
                                                  ctypes.byref(again), None) == 0 and again.value == la.value


# ---------------------------------------------------------------------------------------------------------
# round 2: plans with 4 register bits (sweep-specialised kernels), diagonal tables, phase terms that sink across
# sweeps, and the PTX generator (compiled here with the static PTX compiler, no GPU involved)
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('n,depth,seed,tile', [(12, 20, 0, 12), (12, 6, 1, 8), (11, 5, 2, 11), (9, 4, 3, 7),
                                                (13, 3, 5, 13)])
def test_wb_plan_with_four_register_bits_matches_oracle(n, depth, seed, tile):
    specs = workloads.wb_gate_list(n, depth, seed)
    segments = planner.build_segments(n, bitops_of(specs, n), tile_bits=tile, reg_bits=4)
    assert all(E.parse(s.blob)['sweeps'] is not None for s in segments if s.kind == 'plan')
    assert E.R == 4
    got = run_segments(segments, zero(n))
    assert np.abs(got - O.run_specs(specs, n).reshape(-1)).max() < AMP_TOL


def test_phase_terms_sink_to_their_anchor_and_form_tables():
    """T / RZ / CZ-heavy circuit: phase terms move to the sweep and round of the operator that next mixes their bit
    and are multiplied into diagonal tables there; the result is unchanged."""
    n = 12
    rnd = random.Random(7)
    specs = []
    for layer in range(10):
        for q in range(n):
            kind = rnd.choice(['T', 'RZ', 'S', 'H', 'RX'])
            specs.append((kind, (rnd.uniform(0, 6.28),) if kind in ('RZ', 'RX') else (), (q,)))
        qs = list(range(n))
        rnd.shuffle(qs)
        for a, b in zip(qs[::2], qs[1::2]):
            specs.append((rnd.choice(['CZ', 'CNOT']), (), (a, b)))
    for tile, rb in ((8, 5), (8, 4), (12, 4)):
        segments = planner.build_segments(n, bitops_of(specs, n), tile_bits=tile, reg_bits=rb)
        ntables = sum(1 for s in segments for sw in E.parse(s.blob)['sweeps'] for rd in sw['rounds']
                      for op in rd['ops'] if op['type'] == 4)
        # tables belong to the plans of the sweep-specialised kernels (4 register bits); interpreter plans keep the
        # plain phase placement that the bit-exact sampling fixtures were pinned with
        assert (ntables > 0) == (rb == 4)
        got = run_segments(segments, zero(n))
        assert np.abs(got - O.run_specs(specs, n).reshape(-1)).max() < AMP_TOL


def test_sink_phase_terms_keeps_every_operator_and_moves_terms_forward_only():
    n = 10
    specs = workloads.wb_gate_list(n, 8, 3)
    items = [it for it in planner.classify_all(bitops_of(specs, n)) if not isinstance(it, planner.Fallback)]
    pl = planner.Planner(n, tile_bits=7)
    parts = pl._partition(items)
    sunk = planner.sink_phase_terms(parts)
    before = [id(op) for chosen, _ in parts for op in chosen]
    after = [id(op) for chosen, _ in sunk for op in chosen]
    assert sorted(before) == sorted(after)
    where_before = {id(op): si for si, (chosen, _) in enumerate(parts) for op in chosen}
    where_after = {id(op): si for si, (chosen, _) in enumerate(sunk) for op in chosen}
    assert all(where_after[k] >= where_before[k] for k in where_before)
    # mixing operators never move
    assert [id(op) for chosen, _ in parts for op in chosen if op.kind == 'G'] == \
        [id(op) for chosen, _ in sunk for op in chosen if op.kind == 'G']


@pytest.mark.parametrize('reg_bits', [4, 5])
def test_sweep_specialised_ptx_compiles_for_sm100a(reg_bits):
    """qfb_jit_check: every sweep of a plan -> PTX -> sm_100a image with the statically linked PTX compiler (host
    code only). Checks the generator's output is valid PTX and that the code of a 4-register-bit sweep stays near
    the 32 KiB instruction cache."""
    import ctypes
    from quantumflow_b200 import _lib
    n = 20
    specs = workloads.wb_gate_list(n, 6, 1)
    segments = planner.build_segments(n, bitops_of(specs, n), reg_bits=reg_bits)
    lib = _lib.load()
    for seg in segments:
        log = ctypes.create_string_buffer(1 << 16)
        rc = lib.qfb_jit_check(seg.blob, len(seg.blob), log, len(log))
        assert rc == 0, lib.qfb_last_error()
        text = log.value.decode()
        assert 'Used' in text and 'qfb_sweep' in text
        need, ncoef = ctypes.c_size_t(0), ctypes.c_size_t(0)
        assert lib.qfb_jit_ptx(seg.blob, len(seg.blob), 0, None, 0, ctypes.byref(need), ctypes.byref(ncoef)) == 0
        buf = ctypes.create_string_buffer(need.value)
        assert lib.qfb_jit_ptx(seg.blob, len(seg.blob), 0, buf, need.value, None, None) == 0
        ptx = buf.value.decode()
        assert '.target sm_100a' in ptx and 'cp.async.cg.shared.global' in ptx and 'ld.const.f64' in ptx
        # structure only: no coefficient value appears in the text (one image serves every parameter value)
        assert ptx.count('ld.const.f64') == ncoef.value


def test_latest_possible_rounds_need_no_more_exchanges_and_end_full(monkeypatch):
    """Plans of the sweep-specialised kernels split a sweep's operators into rounds from the back (every operator in
    the latest round that can take it): never more rounds than the forward split, the work sits behind the last
    exchange (where the next tile's asynchronous copy is hidden), and the result is the same."""
    n = 14
    specs = workloads.wb_gate_list(n, 12, 4)
    shapes = {}
    for late in ('0', '1'):
        monkeypatch.setenv('QFB_PLAN_LATE', late)
        segments = planner.build_segments(n, bitops_of(specs, n), tile_bits=11, reg_bits=4)
        sweeps = [sw for s in segments for sw in E.parse(s.blob)['sweeps']]
        shapes[late] = [[len(rd['ops']) for rd in sw['rounds']] for sw in sweeps]
        got = run_segments(segments, zero(n))
        assert np.abs(got - O.run_specs(specs, n).reshape(-1)).max() < AMP_TOL
    assert sum(map(len, shapes['1'])) <= sum(map(len, shapes['0']))
    multi = [s for s in shapes['1'] if len(s) > 1]
    assert multi and sum(s[-1] for s in multi) >= sum(s[0] for s in multi)


@pytest.mark.parametrize('n,depth,seed,reg_bits', [(14, 10, 0, 4), (16, 12, 1, 5), (20, 16, 2, 4)])
def test_native_round_split_matches_the_python_statement(n, depth, seed, reg_bits):
    """qfb_plan_split_rounds (what Planner._form_rounds scores a few hundred variants per sweep with) against
    Planner._split_rounds, the same greedy in Python: plain and randomised variants (the random.Random(trial) stream),
    forward and latest-possible, same operators per round, same register bits, same cost of the last round."""
    import random
    specs = workloads.wb_gate_list(n, depth, seed)
    items = planner.classify_all(bitops_of(specs, n))
    pl = planner.Planner(n, reg_bits=reg_bits)
    sweeps = pl.plan([it for it in items if not isinstance(it, planner.Fallback)])
    checked = 0
    for sw in sweeps:
        ops = [op for rd in sw.rounds for op in rd.ops if op.kind in ('G', 'P')]
        again = planner.SweepPlan(sw.tile, ops)
        splitter = planner._RoundSplitter(pl, again)
        for trial in [0] + list(range(1, 12)):
            for backward in (False, True):
                want = pl._split_rounds(again, random.Random(trial) if trial else None, 0.85 if trial else 1.0, backward)
                nrounds, tail = splitter.score(trial, backward)
                got = splitter.rounds(trial, backward)
                assert nrounds == len(want) == len(got)
                assert tail == sum(op.cost for op in want[-1][1])
                for (wregs, wops), (gregs, gops) in zip(want, got):
                    assert sorted(wregs) == gregs
                    assert [id(op) for op in wops] == [id(op) for op in gops]
                checked += 1
    assert checked >= 24
    # a random stream that is too short is reported, not silently reused
    long_sweep = max(sweeps, key=lambda sw: len(sw.ops))
    splitter = planner._RoundSplitter(pl, planner.SweepPlan(long_sweep.tile, [op for rd in long_sweep.rounds
                                                                           for op in rd.ops if op.kind in ('G', 'P')]))
    saved = planner.ROUND_STREAM_LEN
    try:
        planner.ROUND_STREAM_LEN = 1
        assert splitter.score(3, False) is None
    finally:
        planner.ROUND_STREAM_LEN = saved


def test_tile_search_does_not_depend_on_the_number_of_threads():
    """The candidates of a search pass are scored side by side (csrc/qfb_planhost.cu, PassWorkers); the winner is
    picked in the serial order afterwards, so plans are byte-identical for any thread count."""
    from quantumflow_b200 import _lib
    lib = _lib.load()
    n = 26
    specs = workloads.wb_gate_list(n, 10, 3)
    blobs = {}
    try:
        for threads in (1, 4, 3):
            assert lib.qfb_plan_set_threads(threads) == 0
            segments = planner.build_segments(n, bitops_of(specs, n))
            blobs[threads] = b''.join(s.blob for s in segments if s.blob)
    finally:
        lib.qfb_plan_set_threads(0)
    assert blobs[1] == blobs[4] == blobs[3]

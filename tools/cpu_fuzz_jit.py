#!/usr/bin/env python
"""Time-boxed fuzz of the sweep-specialised kernels WITHOUT a GPU: random circuits (every standard gate kind the
planner has a handler for, random sizes, tiles, register bits, grid sizes, density-style dense 2-bit operators, rank
bits as controls, final bit permutations) -> plan -> generated PTX of every sweep (qfb_jit_source) -> PTX emulator (tests/ptx_emulator.py)
-> compared with the numpy oracle. The GPU twin is tools/gpu_fuzz_jit.py.
Usage: python tools/cpu_fuzz_jit.py [seconds=120] [seed=1]"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                        # noqa: E402
from oracle import qf_oracle as O                         # noqa: E402
from quantumflow_b200 import planner, workloads           # noqa: E402
from test_jit_emulated import run_segments_emulated       # noqa: E402

ONE = ['H', 'X', 'Y', 'Z', 'S', 'T', 'S_H', 'T_H']
ONE_P = ['RX', 'RY', 'RZ', 'TX', 'TY', 'TZ', 'TH', 'PHASE']
TWO = ['CNOT', 'CZ', 'SWAP', 'ISWAP']
TWO_P = ['CPHASE', 'CPHASE00', 'CPHASE01', 'CPHASE10', 'XX', 'YY', 'ZZ', 'PSWAP', 'PISWAP', 'CAN', 'EXCH']
THREE = ['CCNOT', 'CSWAP']

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
cases = bad = sweeps = 0
worst = 0.0
known = {}
while time.time() - t0 < budget:
    n = rnd.randint(9, 13)
    reg_bits = rnd.choice([4, 4, 5])
    tile = rnd.randint(reg_bits + 3, min(n, 11))        # the generator needs at least 3 thread bits
    hi_bits = rnd.choice([0, 0, 1, 2])                 # rank bits: controls / phases only (sharded stages)
    seed = rnd.randint(0, 10 ** 6)
    specs = workloads.wb_gate_list(n, rnd.randint(1, 5), seed) if rnd.random() < 0.5 else []
    ops = [(O.gate_matrix(name, params), [n - 1 - q for q in qubits]) for name, params, qubits in specs]
    for _ in range(rnd.randint(10, 60)):
        family = rnd.choice(['1', '1p', '2', '2p', '3', 'dense2', 'hi'])
        try:
            if family == '1':
                mat, k = O.gate_matrix(rnd.choice(ONE)), 1
            elif family == '1p':
                mat, k = O.gate_matrix(rnd.choice(ONE_P), (rnd.uniform(-3, 3),)), 1
            elif family == '2':
                mat, k = O.gate_matrix(rnd.choice(TWO)), 2
            elif family == '2p':
                name = rnd.choice(TWO_P)
                params = tuple(rnd.uniform(-1, 1) for _ in range(3 if name in ('CAN', 'EXCH') else 1))
                if name == 'EXCH':
                    params = params[:1]
                mat, k = O.gate_matrix(name, params), 2
            elif family == '3':
                mat, k = O.gate_matrix(rnd.choice(THREE)), 3
            elif family == 'dense2':
                rs = np.random.RandomState(rnd.randint(0, 10 ** 6))
                mat, k = rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4)), 2
                mat = mat / np.linalg.norm(mat, 2)
            else:
                if not hi_bits:
                    continue
                # controlled / diagonal operators whose control sits on a rank bit (index_hi resolves it)
                name = rnd.choice(['CNOT', 'CZ', 'CPHASE'])
                mat, k = O.gate_matrix(name, (rnd.uniform(-3, 3),) if name == 'CPHASE' else ()), 2
                bits = [n + rnd.randrange(hi_bits), rnd.randrange(n)]
                ops.append((mat, bits))
                continue
        except Exception as exc:                         # a gate name the oracle does not know: skip the family
            known[str(exc)[:60]] = known.get(str(exc)[:60], 0) + 1
            continue
        bits = rnd.sample(range(n), k)
        ops.append((mat, bits))
    total = n + hi_bits
    rng = np.random.RandomState(seed)
    full = rng.normal(size=1 << total) + 1j * rng.normal(size=1 << total)
    full /= np.linalg.norm(full)
    want = full.copy()
    for mat, bits in ops:
        want = O.tensormul_flat(np.asarray(mat, dtype=np.complex128), want, list(bits))
    # a third of the cases end with an in-place bit permutation of the local index (the local half of a qubit remap:
    # fused into the last sweep's store where its tile allows, bare sweeps otherwise)
    perm = None
    if rnd.random() < 0.33:
        perm = list(range(n))
        if rnd.random() < 0.5:
            movers = rnd.sample(range(3, n), rnd.randint(1, 3))
            perm = [b for b in range(n) if b not in movers] + movers
        else:
            rnd.shuffle(perm)
        idx = np.arange(1 << total)
        src = idx & ~((1 << n) - 1)                       # rank bits stay
        for j in range(n):
            src |= ((idx >> j) & 1) << perm[j]            # destination bit j <- source bit perm[j]
        want = want[src]
    segments = planner.build_segments(n, ops, tile_bits=tile, reg_bits=reg_bits, final_perm=perm)
    got = full.copy()
    for hi in range(1 << hi_bits):
        shard = np.ascontiguousarray(got[hi << n:(hi + 1) << n])
        run_segments_emulated(segments, shard, index_hi=hi, grid=rnd.randint(1, 4))
        got[hi << n:(hi + 1) << n] = shard
    err = float(np.abs(got - want).max())
    sweeps += sum(s.nsweeps for s in segments if s.kind == 'plan')
    worst = max(worst, err)
    cases += 1
    if not err < 1e-10:
        bad += 1
        print('BAD n={} tile={} reg_bits={} hi_bits={} seed={} ops={} err={}'.format(n, tile, reg_bits, hi_bits, seed,
                                                                                len(ops), err), flush=True)
print('cpu fuzz: {} circuits, {} sweeps generated and emulated, {} bad, worst max-abs error {:.2e}; skipped: {}'.format(
    cases, sweeps, bad, worst, known))
sys.exit(1 if bad else 0)

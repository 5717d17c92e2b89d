#!/usr/bin/env python
"""Time-boxed fuzz of the sweep-specialised kernels on a GPU: random layered circuits (plus extra CNOT / CZ / SWAP /
ISWAP / CCNOT) on 10-15 qubits, random tile sizes, kernels forced on (QFB_JIT=1), every result against the numpy
oracle. Usage: python tools/gpu_fuzz_jit.py [seconds=90] [seed=1]"""
import os
import random
import sys
import time

os.environ['QFB_JIT'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                        # noqa: E402
import torch                                              # noqa: E402
from oracle import qf_oracle as O                         # noqa: E402
from quantumflow_b200 import engine, planner, workloads   # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
cases = bad = sweeps = 0
worst = 0.0
while time.time() - t0 < budget:
    n = rnd.randint(10, 15)
    depth = rnd.randint(2, 8)
    seed = rnd.randint(0, 10 ** 6)
    tile = rnd.randint(8, min(n, 12))
    specs = workloads.wb_gate_list(n, depth, seed)
    extra = random.Random(seed)
    for _ in range(extra.randint(0, 8)):
        kind = extra.choice(['CNOT', 'CZ', 'SWAP', 'ISWAP', 'CCNOT'])
        qs = tuple(extra.sample(range(n), 3 if kind == 'CCNOT' else 2))
        specs.insert(extra.randint(0, len(specs)), (kind, (), qs))
    bitops = [(O.gate_matrix(name, params), [n - 1 - q for q in qubits]) for name, params, qubits in specs]
    segments = planner.build_segments(n, bitops, tile_bits=tile, reg_bits=4)
    state = torch.zeros(1 << n, dtype=torch.complex128, device='cuda')
    state[0] = 1.0
    for seg in segments:
        if seg.kind == 'plan':
            engine.UploadedPlan(seg.blob).launch(state)
            sweeps += seg.nsweeps
        else:
            engine.apply_operator(state, seg.mat, seg.bits, inplace=True)
    err = float(np.abs(state.cpu().numpy() - O.run_specs(specs, n).reshape(-1)).max())
    worst = max(worst, err)
    cases += 1
    if not err < 1e-10:
        bad += 1
        print('BAD n={} depth={} seed={} tile={} err={}'.format(n, depth, seed, tile, err))
print('fuzz: {} circuits, {} sweeps compiled and run, {} bad, worst max-abs error {:.2e}'.format(cases, sweeps, bad, worst))
sys.exit(1 if bad else 0)

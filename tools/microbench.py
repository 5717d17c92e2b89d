#!/usr/bin/env python
"""Micro-benchmarks of the sweep kernel on synthetic plans (GPU): the cost of a bare sweep (load + store), of a
round transition, and of each op kind, from differences between plans. Prints one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf                          # noqa: E402
from quantumflow_b200 import engine, planner           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 12


def time_plan(name, bitops, reps=5, **kw):
    segs = planner.build_segments(N, bitops, tile_bits=TILE, max_cost=1e9, **kw)
    assert len(segs) == 1 and segs[0].kind == 'plan'
    up = engine.UploadedPlan(segs[0].blob)
    state = torch.zeros(1 << N, dtype=torch.complex128, device='cuda')
    state[0] = 1
    for _ in range(2):
        up.launch(state)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        up.launch(state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = planner.plan_stats(segs)
    print(json.dumps({'case': name, 'ms': round(ms, 3), 'sweeps': st['sweeps'], 'rounds': st['rounds'],
                      'ops': len(bitops), 'GBps': round(32.0 * (1 << N) * st['sweeps'] / ms / 1e6, 1)}), flush=True)
    del state
    return ms


H = qf.H(0).matrix()
X = qf.X(0).matrix()
T = qf.T(0).matrix()
RX = qf.RX(0.7, 0).matrix()
RY = qf.RY(0.9, 0).matrix()
CNOT = qf.CNOT(0, 1).matrix()
CZ = qf.CZ(0, 1).matrix()
GEN = qf.TX(0.37, 0).matrix()

hi = [20, 21, 22, 23]          # four high bits -> one round
base = time_plan('1 round, 1 scalar phase (bare sweep)', [(T, [25])])
time_plan('1 round, 4 H on 4 register bits', [(H, [b]) for b in hi])
for k in (8, 16, 32, 64):
    time_plan('1 round, %d H' % k, [(H, [hi[i % 4]]) for i in range(k)])
time_plan('1 round, 32 RX', [(RX, [hi[i % 4]]) for i in range(32)])
time_plan('1 round, 32 RY', [(RY, [hi[i % 4]]) for i in range(32)])
time_plan('1 round, 32 general', [(GEN, [hi[i % 4]]) for i in range(32)])
time_plan('1 round, 32 X + 32 H on the same bits (X = flip mask: should cost like 32 H)', [((X if i % 2 else H), [hi[(i // 2) % 4]]) for i in range(64)])
time_plan('1 round, 32 CNOT reg-reg', [(CNOT, [hi[i % 4], hi[(i + 1) % 4]]) for i in range(32)])
time_plan('1 round, 32 CNOT thread-ctrl (bit 26 -> reg)', [(CNOT, [26, hi[i % 4]]) for i in range(32)])
time_plan('1 round, 32 CNOT lane-ctrl (bit 1 -> reg)', [(CNOT, [1, hi[i % 4]]) for i in range(32)])
time_plan('1 round, 32 T on reg bits (+H between, same bit)', [((T if i % 2 else H), [hi[(i // 2) % 4]]) for i in range(64)])
time_plan('1 round, 32 CZ reg-reg (+H between)', [((CZ, [hi[i % 4], hi[(i + 1) % 4]]) if i % 2 else (H, [hi[i % 4]])) for i in range(64)])
time_plan('32 scalar phases on distinct thread bits', [(qf.RZ(0.1 * i, 0).matrix(), [4 + (i % 12)]) for i in range(32)])
# rounds: H on 4r distinct bits
for r in (2, 3, 4, 6):
    bits = list(range(3, 3 + 4 * r)) if 3 + 4 * r <= 3 + 9 else None
    if bits is None:
        bits = [3 + (i % 9) for i in range(4 * r)]
    # force separate rounds by chaining dependencies through distinct bit groups
    time_plan('%d rounds x 4 H' % r, [(H, [b]) for b in bits])

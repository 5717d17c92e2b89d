#!/usr/bin/env python
"""Where does the time of the 30-qubit benchmark plan go? Times the real plan and stripped variants of it
(same tiles and rounds without ops; same tiles with one empty round; contiguous tiles) on the GPU."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf                               # noqa: E402
from quantumflow_b200 import engine, planner, workloads     # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 12
LOW = int(sys.argv[3]) if len(sys.argv) > 3 else 3

specs = workloads.wb_gate_list(N, 20, 0)
bitops = [(g.matrix(), [N - 1 - q for q in g.qubits]) for g in workloads.circuit_from_specs(qf, specs).elements]
P = planner.Planner(N, tile_bits=TILE, low_bits=LOW)
pops = []
for gi, (m, b) in enumerate(bitops):
    it = planner.classify_op(np.asarray(m), list(b), gi)
    if it:
        pops.extend(it)
sweeps = P.plan(pops)
state = torch.zeros(1 << N, dtype=torch.complex128, device='cuda')
state[0] = 1


def run(name, sws, reps=3):
    blob = P.serialise(sws)
    up = engine.UploadedPlan(blob)
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
    print(json.dumps({'case': name, 'ms': round(ms, 2), 'sweeps': len(sws), 'ms_per_sweep': round(ms / len(sws), 3),
                      'rounds': sum(len(s.rounds) for s in sws), 'ops': sum(len(r.ops) for s in sws for r in s.rounds)}),
          flush=True)


def strip(sw, keep_rounds, kinds=()):
    s2 = planner.SweepPlan(list(sw.tile), [], sw.store_xor)
    rounds = sw.rounds if keep_rounds else [sw.rounds[0]] if len(sw.rounds) == 1 else [sw.rounds[0]]
    s2.rounds = []
    for rd in rounds:
        ops = [op for op in rd.ops if ('P' in kinds and op.kind == 'P') or
               (op.kind == 'G' and (('C' in kinds and op.ctrl) or ('G' in kinds and not op.ctrl)))]
        s2.rounds.append(planner.Round(list(rd.regs), list(rd.thr), ops))
    if not keep_rounds:
        # one round that both loads and stores: register bits must avoid the low tile positions
        rd = s2.rounds[0]
        if any(p < P.L for p in rd.regs):
            regs = [p for p in range(P.M - 1, -1, -1) if p >= P.L][:planner.REG_BITS]
            regs.sort()
            s2.rounds = [planner.Round(regs, P._thread_order(regs, True), [])]
        s2.store_xor = 0
    return s2


run('full plan', sweeps)
run('same tiles + rounds, no ops', [strip(s, True) for s in sweeps])
run('same tiles + rounds, phase terms only', [strip(s, True, ('P',)) for s in sweeps])
run('same tiles + rounds, uncontrolled 1-bit ops only', [strip(s, True, ('G',)) for s in sweeps])
run('same tiles + rounds, controlled ops only', [strip(s, True, ('C',)) for s in sweeps])
run('same tiles, one empty round', [strip(s, False) for s in sweeps])
contig = []
for s in sweeps:
    c = planner.SweepPlan(list(range(P.M)), [], 0)
    regs = list(range(P.M - planner.REG_BITS, P.M))
    c.rounds = [planner.Round(regs, P._thread_order(regs, True), [])]
    contig.append(c)
run('contiguous tiles, one empty round', contig)

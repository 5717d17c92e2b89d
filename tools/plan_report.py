#!/usr/bin/env python
"""Composition of the plan of a W-B benchmark circuit (CPU only, no GPU): sweeps, rounds, operators per sweep and
the handler-level histogram (which kinds of operators the sweep kernel will interpret, how many of their bits sit
in registers). Usage: python tools/plan_report.py [qubits=30] [depth=20] [seed=0]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf                      # noqa: E402
from quantumflow_b200 import planner, workloads    # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
circ = workloads.wb_circuit(qf, n, depth, seed)
bitops = [(g.matrix(), [n - 1 - circ.qubits.index(q) for q in g.qubits]) for g in circ.elements]
items = planner.classify_all(bitops)
t0 = time.perf_counter()
sweeps = planner.Planner(n).plan([it for it in items if not isinstance(it, planner.Fallback)])
print('{} gates -> {} classified operators -> {} sweeps, {} rounds ({:.2f} s)'.format(
    len(bitops), len(items), len(sweeps), sum(len(s.rounds) for s in sweeps), time.perf_counter() - t0))
hist = collections.Counter()
for s in sweeps:
    pos_of = {b: j for j, b in enumerate(s.tile)}
    for rd in s.rounds:
        for op in rd.ops:
            if op.kind == 'G':
                in_regs = sum(1 for b in op.ctrl if pos_of.get(b) in rd.regs)
                key = ('G', 'kind %s' % (op.enc[0] if op.enc else '-'), '%d controls' % len(op.ctrl),
                       '%d in registers' % in_regs)
            else:
                in_regs = sum(1 for b in op.dbits if pos_of.get(b) in rd.regs)
                key = ('P', '%d bits' % len(op.dbits), '%d in registers' % in_regs, 'sign' if op.mat == -1 else 'phase')
            hist[key] += 1
print('operators per sweep (rounds):', ' '.join('{}({})'.format(len(s.ops), len(s.rounds)) for s in sweeps))
for key, count in sorted(hist.items(), key=lambda kv: -kv[1]):
    print('{:5d}  {}'.format(count, ', '.join(key)))

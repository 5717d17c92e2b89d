#!/usr/bin/env python
"""FP64-pipe and integer-pipe instruction estimate of a plan, per thread and tile (CPU only). The FP64 pipe of a
B200 SM issues 1.83 warp instructions per clock (profiles/r2_fp64_peaks.jsonl: 34 TFLOP/s DFMA; DMMA shares the
pipe), so a 30-qubit sweep that has F FP64 instructions per thread keeps it busy for F * 2^20 / 5.3e11 seconds.
Usage: python tools/plan_cost.py [qubits=30] [depth=20] [seed=0] [tile_bits=12]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import plan_emulator as E                               # noqa: E402


def popcount(x):
    return bin(x).count('1')


def op_cost(op, handler_names=None):
    """(fp64, alu) instructions per thread of one decoded op (E.NE amplitudes per thread: 32 or 16)."""
    rc = popcount(op['reg_cmask'])
    ne = E.NE
    if op['type'] == 1:
        pairs = (ne // 2) >> rc
        k = op['kind']
        if k == 'general':
            return 16 * pairs, 0
        if k == 'swapx':
            return (0, 0) if op.get('free') else (6 * pairs, 0)
        return 4 * pairs, 0
    if op['type'] == 2:
        groups = (ne // 4) >> rc
        return (16 * groups if op['kind'] == 'xshape' else 128 * groups), 0
    if op['type'] == 4:     # diagonal table
        return (4 * ne if op['flag'] else 4 * (ne - (ne >> popcount(op['reg_cmask'])))), 0
    # phase terms
    if op['reg_cmask'] == 0:
        return 4, 0
    touched = ne >> rc
    if op['kind'] == 'neg':
        return 0, 2 * touched
    return (2 if op.get('real') else 4) * touched, 0


def plan_cost(blob):
    plan = E.parse(blob)
    rows = []
    for sw in plan['sweeps']:
        fp = alu = 0
        for rd in sw['rounds']:
            for op in rd['ops']:
                f, a = op_cost(op)
                fp += f
                alu += a
            if rd['has_scalar']:
                fp += 4 * E.NE
        rows.append((len(sw['rounds']), sum(len(rd['ops']) for rd in sw['rounds']), fp, alu))
    return rows


if __name__ == '__main__':
    import numpy as np
    import quantumflow_b200 as qf
    from quantumflow_b200 import planner, workloads
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    tile = int(sys.argv[4]) if len(sys.argv) > 4 else None
    circ = workloads.wb_circuit(qf, n, depth, seed)
    bitops = [(g.matrix(), [n - 1 - circ.qubits.index(q) for q in g.qubits]) for g in circ.elements]
    segs = planner.build_segments(n, bitops, tile_bits=tile)
    total_fp = total_alu = 0
    for seg in segs:
        if seg.kind != 'plan':
            continue
        for i, (nr, nops, fp, alu) in enumerate(plan_cost(seg.blob)):
            print('sweep {:2d}: {} rounds {:3d} ops  fp64/thread {:5d}  alu/thread {:5d}  fp64-pipe ms {:.2f}'.format(
                i, nr, nops, fp, alu, fp * (1 << (n - 10)) / 5.3e11 * 1e3))
            total_fp += fp
            total_alu += alu
    print('total fp64/thread {}  alu {}  fp64-pipe floor {:.1f} ms'.format(total_fp, total_alu,
                                                                        total_fp * (1 << (n - 10)) / 5.3e11 * 1e3))

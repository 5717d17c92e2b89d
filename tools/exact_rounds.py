#!/usr/bin/env python
"""Is the planner's split of a sweep into rounds optimal? (CPU only.) For every sweep of the W-B benchmark plan that
needs 4 rounds, enumerate the register sets of a 3-round split exhaustively (first and last round: 4 of the tile
positions above the low bits, middle round: any 4 positions; for fixed register sets the greedy earliest-round
assignment of Planner._split_rounds is optimal, because a round has no capacity other than its register bits) and
report whether a 3-round split exists. Usage: python tools/exact_rounds.py [qubits=30] [depth=20] [seed=0]"""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qf_oracle as O                        # noqa: E402
from quantumflow_b200 import planner, workloads          # noqa: E402


def three_round_split(pl, tile, ops):
    """(register mask round 1, round 2, round 3) of a 3-round split, or None when none exists."""
    M, L, R = len(tile), pl.L, pl.R
    pos_of = {b: j for j, b in enumerate(tile)}
    info = [(op.mixmask, op.diagmask, sum(1 << pos_of[b] for b in op.mix) if op.kind == 'G' else None) for op in ops]
    low = (1 << L) - 1

    def leftover(remaining, regs):
        da = dm = 0
        rest = []
        for i in remaining:
            mm, dd, pm = info[i]
            if (mm & da) or (dd & dm) or (pm is not None and (pm & ~regs)):
                rest.append(i)
                da |= mm | dd
                dm |= mm
        return rest

    first = {}
    for c1 in itertools.combinations(range(L, M), R):
        r1 = sum(1 << p for p in c1)
        first.setdefault(tuple(leftover(range(len(ops)), r1)), r1)
    for rest1, r1 in first.items():
        seen = set()
        for c2 in itertools.combinations(range(M), R):
            r2 = sum(1 << p for p in c2)
            rest2 = tuple(leftover(rest1, r2))
            if rest2 in seen:
                continue
            seen.add(rest2)
            need = 0
            for i in rest2:
                if info[i][2] is not None:
                    need |= info[i][2]
            if not (need & low) and bin(need).count('1') <= R:
                return r1, r2, need
    return None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    specs = workloads.wb_gate_list(n, depth, seed)
    bitops = [(O.gate_matrix(name, params), [n - 1 - q for q in qubits]) for name, params, qubits in specs]
    captured = []
    original = planner.Planner._form_rounds

    def hook(self, sweep):
        ops = list(sweep.ops)
        original(self, sweep)
        captured.append((self, list(sweep.tile), ops, len(sweep.rounds)))

    planner.Planner._form_rounds = hook
    try:
        planner.build_segments(n, bitops)
    finally:
        planner.Planner._form_rounds = original
    print('rounds per sweep:', [c[3] for c in captured], '=', sum(c[3] for c in captured))
    for idx, (pl, tile, ops, nrounds) in enumerate(captured):
        if nrounds >= 4:
            t0 = time.perf_counter()
            found = three_round_split(pl, tile, ops)
            print('sweep {:2d}: {} operators, {} rounds; a 3-round split {} ({:.1f} s)'.format(
                idx, len(ops), nrounds, 'EXISTS: register masks {}'.format(found) if found else 'does not exist',
                time.perf_counter() - t0))


if __name__ == '__main__':
    main()

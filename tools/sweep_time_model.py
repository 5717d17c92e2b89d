#!/usr/bin/env python
"""What explains the duration of a benchmark sweep? (CPU only: the committed ncu launch list + the plan.)
Per sweep of the 30-qubit benchmark plan: rounds, static instruction counts of the generated tile loop (FP64, selects,
shared-memory accesses), the tile's index bits, and the measured duration (profiles/r2_launches_final.csv, two steps).
Least-squares fits of the duration on these features, printed with their residuals.
Usage: python tools/sweep_time_model.py [launch list csv]"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                        # noqa: E402
import plan_emulator as E                                 # noqa: E402
from oracle import qf_oracle as O                         # noqa: E402
from quantumflow_b200 import planner, workloads           # noqa: E402
from test_jit_emulated import sweep_source                # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r2_launches_final.csv')
n = 30
specs = workloads.wb_gate_list(n, 20, 0)
segments = planner.build_segments(n, [(O.gate_matrix(a, b), [n - 1 - q for q in c]) for a, b, c in specs])
blob = segments[0].blob
plan = E.parse(blob)
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
header = rows[0]
ki, vi = header.index('Kernel Name'), header.index('Metric Value')
times = [float(r[vi].replace(',', '')) / 1e6 for r in rows[1:] if 'qfb_sweep' in r[ki]]
ns = len(plan['sweeps'])
assert len(times) >= 2 * ns, 'the launch list does not hold two steps of this plan'
t = (np.array(times[:ns]) + np.array(times[ns:2 * ns])) / 2
features = []
for i, sw in enumerate(plan['sweeps']):
    ptx = sweep_source(blob, i)[0]
    count = collections.Counter()
    for line in ptx[ptx.index('L_TILE:'):].splitlines():
        line = re.sub(r'^@!?%p\d+\s+', '', line.strip())
        if line and not line.endswith(':'):
            count[line.split()[0].rstrip(';')] += 1
    fp64 = count['fma.rn.f64'] + count['mul.f64']
    features.append((len(sw['rounds']), fp64, count['selp.f64'], count['ld.shared.v2.f64'] + count['st.shared.v2.f64'],
                     sum(count.values())))
    print('sweep {:2d}: {:.2f} ms  rounds {}  fp64 {:4d}  selp {:3d}  lds/sts {:3d}  instructions {:4d}  tile bits {}'
          .format(i, t[i], features[-1][0], fp64, features[-1][2], features[-1][3], features[-1][4], sw['gpos']))
F = np.array(features, dtype=float)
tiles = [sw['gpos'] for sw in plan['sweeps']]


def has(bit):
    return np.array([1.0 if bit in g else 0.0 for g in tiles])


def fit(name, columns):
    A = np.column_stack([np.ones(ns)] + columns)
    coef = np.linalg.lstsq(A, t, rcond=None)[0]
    res = A @ coef - t
    print('{:34s} coefficients {}  rms residual {:.2f} ms  R^2 {:.2f}'.format(
        name, np.round(coef, 3), float(np.sqrt(np.mean(res ** 2))), 1 - float(np.sum(res ** 2) / np.sum((t - t.mean()) ** 2))))


print()
fit('FP64 instructions', [F[:, 1]])
fit('all instructions', [F[:, 4]])
fit('rounds', [F[:, 0]])
fit('rounds + FP64', [F[:, 0], F[:, 1]])
fit('rounds + FP64 + bit 3 in tile', [F[:, 0], F[:, 1], has(3)])
fit('rounds + FP64 + bit 3 + bit 5', [F[:, 0], F[:, 1], has(3), has(5)])
fit('rounds + (bit 3 or bit 5 in tile)', [F[:, 0], np.maximum(has(3), has(5))])

#!/usr/bin/env python
"""Wall-clock of the BASELINE.json configurations that are parity cases rather than bench lines (C1 20-qubit random
circuit, C3 14-qubit density evolution with depolarizing channels), through the public API on one GPU."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf                      # noqa: E402
from quantumflow_b200 import engine, workloads     # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


c1 = workloads.wb_circuit(qf, 20, 20, 0)
ms = timed(lambda: c1.run())
print(json.dumps({'config': 'C1: W-B 20 qubits depth 20 (620 gates), Circuit.run incl. zero-state creation',
                  'ms': round(ms, 3), 'gates_per_s': round(620 / ms * 1e3)}))
c3 = workloads.wd_circuit(qf, 14, 20, 0, kraus=True)
n_ops = len(c3.elements)
ms = timed(lambda: c3.evolve(), reps=3)
print(json.dumps({'config': 'C3: W-D density 14 qubits depth 20 ({} ops: RX, CNOT, Depolarizing Kraus), '
                            'Circuit.evolve'.format(n_ops), 'ms': round(ms, 3), 'ops_per_s': round(n_ops / ms * 1e3)}))
c3b = workloads.wd_circuit(qf, 14, 20, 0, kraus=False)
ms = timed(lambda: c3b.evolve(), reps=3)
print(json.dumps({'config': 'C3 (channels as superoperators)', 'ms': round(ms, 3),
                  'ops_per_s': round(n_ops / ms * 1e3)}))

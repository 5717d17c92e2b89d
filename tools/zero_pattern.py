#!/usr/bin/env python
"""Exact-zero pattern of Circuit.run against the oracle (the reference's own einsum arithmetic) for the circuit of the
bit-exact sampling test, under the planner's experiment toggles (GPU). Sampling with np.random.multinomial skips
zero-probability bins without consuming random numbers, so the PATTERN of exact zeros must equal the reference's."""
import itertools, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf
from quantumflow_b200 import workloads
from oracle import qf_oracle as O

n, depth, seed = 8, 3, 9
specs = workloads.wb_gate_list(n, depth, seed)
want = O.run_specs(specs, n).reshape(-1)
names = ['QFB_PLAN_SINK', 'QFB_PLAN_TABLES', 'QFB_PLAN_CARRY', 'QFB_PLAN_ALLIN']
for combo in itertools.product('01', repeat=4):
    for k, v in zip(names, combo):
        os.environ[k] = v
    got = qf.asarray(workloads.wb_circuit(qf, n, depth, seed).run().tensor).reshape(-1)
    extra = (want == 0) & (got != 0)
    missing = (want != 0) & (got == 0)
    print(dict(zip([s[9:] for s in names], combo)), 'spurious nonzeros', int(extra.sum()), 'max', float(np.abs(got[extra]).max()) if extra.any() else 0.0,
          'spurious zeros', int(missing.sum()), 'maxerr %.2e' % np.abs(got - want).max())

"""Golden fixture for batched trajectories (tests/golden/trajectories.npz), produced by RUNNING THE REFERENCE in the
build container:  python tests/golden/make_golden_trajectories.py
The reference unravels channels one state at a time (quantumflow/channels.py:70-77, 119-125); the fixture is the
loop `for op in circuit: for t in range(B): ket[t] = op.run(ket[t])` under a fixed numpy seed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from refshim import load_reference   # noqa: E402

qf = load_reference()
bk = qf.backend


def build(qfmod, n):
    """The same operations with any module offering the QuantumFlow API (also imported by the GPU test)."""
    ops = [qfmod.H(q) for q in range(n)]
    ops += [qfmod.CNOT(q, q + 1) for q in range(0, n - 1, 2)]
    ops += [qfmod.Damping(0.3, q) for q in range(n)]
    ops += [qfmod.RX(0.4 + 0.1 * q, q) for q in range(n)]
    ops += [qfmod.Depolarizing(0.4, q) for q in range(n)]
    ops += [qfmod.CZ(q, q + 1) for q in range(1, n - 1, 2)]
    ops += [qfmod.RY(1.1 - 0.2 * q, q) for q in range(n)]
    ops += [qfmod.Dephasing(0.5, q) for q in range(n)]
    ops += [qfmod.Damping(0.6, q) for q in (0, n - 1)]
    return ops


if __name__ == '__main__':
    out = {}
    for name, n, batch, seed in (('a', 5, 8, 7), ('b', 7, 16, 11)):
        np.random.seed(seed)
        kets = [qf.zero_state(n) for _ in range(batch)]
        for op in build(qf, n):
            kets = [op.run(k) for k in kets]
        out[name + '_kets'] = np.stack([np.asarray(bk.evaluate(k.tensor)).reshape(-1) for k in kets])
        out[name + '_next_uniform'] = np.asarray([np.random.random_sample()])
        out[name + '_meta'] = np.asarray([n, batch, seed])
    np.savez_compressed(os.path.join(HERE, 'trajectories.npz'), **out)
    print('wrote trajectories.npz', {k: v.shape for k, v in out.items()})

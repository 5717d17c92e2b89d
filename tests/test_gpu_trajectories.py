"""Batched stochastic trajectories on the device against the REFERENCE's one-state-at-a-time
unravelling (quantumflow/channels.py:70-77, 119-125) under the shared numpy RNG stream: fixture produced by
tests/golden/make_golden_trajectories.py, which runs the reference loop
`for op in circuit: for t in range(B): ket[t] = op.run(ket[t])`."""
import os
import sys

import numpy as np
import pytest
import torch

import quantumflow_b200 as qf
from quantumflow_b200 import engine
from quantumflow_b200.trajectories import StateBatch

from conftest import AMP_TOL

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))


def _build(n):
    # the same operation list as make_golden_trajectories.build (not imported: that module loads the reference)
    ops = [qf.H(q) for q in range(n)]
    ops += [qf.CNOT(q, q + 1) for q in range(0, n - 1, 2)]
    ops += [qf.Damping(0.3, q) for q in range(n)]
    ops += [qf.RX(0.4 + 0.1 * q, q) for q in range(n)]
    ops += [qf.Depolarizing(0.4, q) for q in range(n)]
    ops += [qf.CZ(q, q + 1) for q in range(1, n - 1, 2)]
    ops += [qf.RY(1.1 - 0.2 * q, q) for q in range(n)]
    ops += [qf.Dephasing(0.5, q) for q in range(n)]
    ops += [qf.Damping(0.6, q) for q in (0, n - 1)]
    return ops


@pytest.mark.parametrize('name', ['a', 'b'])
def test_batched_trajectories_match_the_reference_loop(golden, name):
    data = golden('trajectories.npz')
    n, batch, seed = (int(v) for v in data[name + '_meta'])
    np.random.seed(seed)
    before = engine.launch_count()
    out = StateBatch.zeros(batch, n).run(qf.Circuit(_build(n)))
    got = out.asarray()
    # amplitudes of every trajectory (same branches drawn) and the position of the shared RNG stream afterwards
    assert np.abs(got - data[name + '_kets']).max() < AMP_TOL
    assert np.random.random_sample() == float(data[name + '_next_uniform'][0])
    assert np.abs(out.norms() - 1).max() < 1e-12
    assert engine.launch_count() > before


def test_batch_agrees_with_per_state_runs_on_random_states():
    """Random normalised input states, 10 qubits x 32 trajectories: the batch equals Kraus.run / Gate.run applied state
    by state in the reference's loop order with the same seed (the per-state path is itself pinned to the
    reference, tests/test_gpu_states.py)."""
    n, batch = 10, 32
    rng = np.random.RandomState(3)
    vecs = rng.normal(size=(batch, 1 << n)) + 1j * rng.normal(size=(batch, 1 << n))
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    states = [qf.State(v.reshape([2] * n)) for v in vecs]
    ops = _build(n)
    np.random.seed(21)
    seq = list(states)
    for op in ops:
        seq = [op.run(k) for k in seq]
    want = np.stack([qf.asarray(k.tensor).reshape(-1) for k in seq])
    np.random.seed(21)
    got = StateBatch.from_states(states).run(qf.Circuit(ops)).asarray()
    assert np.abs(got - want).max() < AMP_TOL


def test_reduced_density_and_per_trajectory_operator_kernels():
    """qfb_batch_rho1 / qfb_batch_apply1 against numpy on every bit position."""
    n, b = 9, 3
    rng = np.random.RandomState(8)
    vecs = rng.normal(size=(1 << b, 1 << n)) + 1j * rng.normal(size=(1 << b, 1 << n))
    batch = StateBatch(torch.from_numpy(vecs.copy()).cuda(), tuple(range(n)), b)
    for bit in range(n):
        rho = batch._reduced_density(bit)
        cube = vecs.reshape(1 << b, 1 << (n - 1 - bit), 2, 1 << bit)
        x, y = cube[:, :, 0, :], cube[:, :, 1, :]
        want = np.stack([(np.abs(x) ** 2).sum(axis=(1, 2)), (np.abs(y) ** 2).sum(axis=(1, 2)),
                         (x * y.conj()).sum(axis=(1, 2)).real, (x * y.conj()).sum(axis=(1, 2)).imag], axis=1)
        assert np.abs(rho - want).max() < 1e-10
    mats = rng.normal(size=(1 << b, 2, 2)) + 1j * rng.normal(size=(1 << b, 2, 2))
    bit = 4
    batch._apply_per_trajectory(bit, mats)
    cube = vecs.reshape(1 << b, 1 << (n - 1 - bit), 2, 1 << bit)
    want = np.einsum('tij,tajb->taib', mats, cube).reshape(1 << b, -1)
    assert np.abs(batch.asarray() - want).max() < 1e-12

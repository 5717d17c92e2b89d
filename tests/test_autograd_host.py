"""Host half of the single-launch circuit node (quantumflow_b200/autograd.py): eligibility and the vectorised
unitarity test. The kernels themselves are covered by tests/test_gpu_autograd.py."""
import numpy as np
import torch

import quantumflow_b200 as qf
from quantumflow_b200 import autograd


def test_shape_rule_takes_one_and_two_qubit_gates_on_small_states():
    gates = [qf.H(0), qf.CNOT(0, 1), qf.RX(0.3, 2)]
    assert autograd.small_circuit_shape_ok(gates, 3)
    assert autograd.small_circuit_shape_ok(gates, autograd.SMALL_MAX_BITS)
    assert not autograd.small_circuit_shape_ok(gates, autograd.SMALL_MAX_BITS + 1)      # two vectors in shared memory
    assert not autograd.small_circuit_shape_ok(gates[:1], 3)                             # a single gate is not a run
    assert not autograd.small_circuit_shape_ok(gates + [qf.CCNOT(0, 1, 2)], 3)           # 3-qubit gates: gate by gate


def test_vectorised_unitarity_test_finds_the_one_bad_matrix():
    theta = torch.tensor(0.7, dtype=torch.float64, requires_grad=True)
    gates = [qf.H(0), qf.RX(theta, 1), qf.CNOT(0, 1), qf.ZZ(theta, 1, 2), qf.T(2)]
    ks = [g.qubit_nb for g in gates]
    flat = np.concatenate([g.matrix().reshape(-1) for g in gates])
    assert autograd._all_unitary(flat, ks)
    for bad in (qf.P0(0), qf.Gate(np.asarray([[1, 0], [0, 0.5]]), [1])):
        mixed = gates[:2] + [bad] + gates[2:]
        flat = np.concatenate([g.matrix().reshape(-1) for g in mixed])
        assert not autograd._all_unitary(flat, [g.qubit_nb for g in mixed])
    two = qf.Gate(np.diag([1, 1, 1, 2.0]), [0, 1])
    flat = np.concatenate([g.matrix().reshape(-1) for g in gates + [two]])
    assert not autograd._all_unitary(flat, ks + [2])

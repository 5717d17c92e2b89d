"""Host-side operator algebra of quantumflow_b200 (gates, channels, circuits-as-gates) against the
reference-generated fixture tests/golden/stdgates.npz. No GPU involved: operators are planner inputs."""
import json
import os

import numpy as np
import pytest

import quantumflow_b200 as qf
from quantumflow_b200 import classify

from conftest import GOLDEN

PARAMS = json.load(open(os.path.join(GOLDEN, 'stdgates_params.json')))


def mat(gate):
    return qf.asarray(gate.asoperator())


@pytest.mark.parametrize('name', sorted(PARAMS))
def test_gate_matrix_inverse_and_power_match_reference(golden, name):
    data = golden('stdgates.npz')
    gate = qf.STDGATES[name](*PARAMS[name])
    assert np.abs(mat(gate) - data[name]).max() < 1e-15
    assert np.abs(mat(gate.H) - data[name + '__H']).max() < 1e-15
    assert np.abs(mat(gate ** 0.3) - data[name + '__pow']).max() < 1e-12
    assert gate.qubits == tuple(range(gate.qubit_nb))


def test_constructors_match_reference(golden):
    data = golden('stdgates.npz')
    assert np.abs(mat(qf.control_gate(5, qf.RX(0.4, 2))) - data['control_gate_RX']).max() < 1e-15
    assert np.abs(mat(qf.conditional_gate(0, qf.X(1), qf.RY(0.3, 1))) - data['conditional_gate']).max() < 1e-15
    assert np.abs(mat(qf.join_gates(qf.H(0), qf.CNOT(1, 2))) - data['join_gates']).max() < 1e-15
    for key, chan in [('aschannel_RX', qf.RX(0.9, 0).aschannel()), ('aschannel_CNOT', qf.CNOT(0, 1).aschannel()),
                      ('depolarizing_superop', qf.Depolarizing(0.1, 0).aschannel()),
                      ('damping_superop', qf.Damping(0.2, 0).aschannel()),
                      ('dephasing_superop', qf.Dephasing(0.3, 0).aschannel())]:
        got = qf.asarray(chan.tensor).reshape(data[key].shape)
        assert np.abs(got - data[key]).max() < 1e-15, key
    assert np.abs(qf.asarray(qf.Damping(0.2, 0).aschannel().choi()) - data['damping_choi']).max() < 1e-15
    assert np.abs(qf.Depolarizing(0.1, 0).superoperator_matrix() - data['depolarizing_superop']).max() < 1e-15


def test_identities_from_the_reference_tests(golden):
    # 3 CNOTs = SWAP (tests/test_stdgates.py:75-82)
    swap = qf.Circuit([qf.CNOT(0, 1), qf.CNOT(1, 0), qf.CNOT(0, 1)]).asgate()
    assert qf.gates_close(swap, qf.SWAP(0, 1))
    # RZ RX RZ = H up to phase (tests/test_stdgates.py:152-160)
    h = qf.Circuit([qf.RZ(np.pi / 2, 0), qf.RX(np.pi / 2, 0), qf.RZ(np.pi / 2, 0)]).asgate()
    assert qf.gates_close(h, qf.H(0))
    # ZYZ circuit == ZYZ gate (tests/test_circuits.py:27-30)
    assert qf.gates_close(qf.zyz_circuit(0.1, 2.2, 0.5, 0).asgate(), qf.ZYZ(0.1, 2.2, 0.5))
    # CCNOT decomposition (tests/test_circuits.py:234-260)
    assert qf.gates_close(qf.ccnot_circuit([0, 1, 2]).asgate(), qf.CCNOT(0, 1, 2))
    assert np.abs(mat(qf.ccnot_circuit([0, 1, 2]).asgate()) - golden('workloads.npz')['ccnot_circuit_gate']).max() \
        < 1e-14
    # every standard gate is unitary and H is its inverse
    for name, p in PARAMS.items():
        gate = qf.STDGATES[name](*p)
        assert qf.almost_unitary(gate), name
        assert qf.almost_identity(gate.H @ gate), name
    # projectors are Hermitian, not unitary
    assert qf.almost_hermitian(qf.P0()) and not qf.almost_unitary(qf.P1())
    assert qf.kraus_iscomplete(qf.Damping(0.1, 0)) and qf.kraus_iscomplete(qf.Depolarizing(0.2, 0))


def test_gate_api_details():
    g = qf.RX(0.5, 'a')
    assert g.qubits == ('a',) and g.params == {'theta': 0.5} and g.name == 'RX'
    assert qf.TX(2.5).params['t'] == 0.5                      # reduced mod 2 (stdgates.py:685)
    assert isinstance(qf.X(1) ** 0.5, qf.TX) and isinstance(qf.S(0).H, qf.S_H)
    assert qf.I(0, 1, 2).qubit_nb == 3 and qf.identity_gate(2).qubits == (0, 1)
    assert qf.CNOT(2, 3).relabel([7, 8]).qubits == (7, 8)
    perm = qf.CNOT(0, 1).permute([1, 0])
    assert np.array_equal(mat(perm).real, np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]]))
    with pytest.raises(ValueError):
        qf.control_gate(0, qf.X(0))
    with pytest.raises(NotImplementedError):
        qf.X(0) @ 3
    with pytest.raises(TypeError):
        qf.X(0).aschannel().asgate()
    with pytest.raises(ValueError):
        qf.Gate(np.eye(4), qubits=[0, 1, 2])
    assert str(qf.CNOT(0, 1)) == 'CNOT 0 1'
    circ = qf.Circuit([qf.H(3), qf.CNOT(3, 1)])
    assert circ.qubits == (1, 3) and circ.size() == 2 and circ.H.elements[0].name == 'CNOT'
    assert qf.count_operations(circ) == {qf.H: 1, qf.CNOT: 1}


def test_channel_algebra():
    # the transpose map (Kraus form with a negative weight) has SWAP as its Choi matrix
    # (reference tests/test_channels.py:24-47)
    ops = [qf.Gate(np.asarray([[1, 0], [0, 0]])), qf.Gate(np.asarray([[0, 0], [0, 1]])),
           qf.Gate(np.asarray([[0, 1], [1, 0]]) / np.sqrt(2)), qf.Gate(np.asarray([[0, 1], [-1, 0]]) / np.sqrt(2))]
    kraus = qf.Kraus(ops, weights=(1, 1, 1, -1))
    assert np.allclose(qf.asarray(kraus.aschannel().choi()), qf.asarray(qf.SWAP(0, 2).asoperator()))
    a, b = qf.X(0).aschannel(), qf.H(0).aschannel()
    assert qf.channels_close(a @ b, (qf.X(0) @ qf.H(0)).aschannel())
    mix = a * 0.25 + b * 0.75
    want = 0.25 * np.kron(qf.X().matrix(), qf.X().matrix()) + 0.75 * np.kron(qf.H().matrix(), qf.H().matrix())
    assert np.allclose(qf.asarray(mix.tensor).reshape(4, 4), want)
    with pytest.raises(ValueError):
        a + qf.X(1).aschannel()
    kraus = qf.channel_to_kraus(qf.Damping(0.3, 0).aschannel())
    assert qf.channels_close(kraus.aschannel(), qf.Damping(0.3, 0).aschannel())
    assert qf.channels_close(qf.Kraus([qf.X(0)]).H.aschannel(), qf.X(0).aschannel())
    assert qf.Kraus([qf.X(2), qf.CNOT(0, 1)]).qubits == (0, 1, 2)


def test_dagcircuit_depth_and_layers():
    # tests/test_dagcircuit.py:97-113: QFT-4 depth 8, GHZ-5 depth 5 (4 non-local)
    assert qf.DAGCircuit(qf.qft_circuit([0, 1, 2, 3])).depth() == 8
    ghz = qf.DAGCircuit(qf.ghz_circuit(range(5)))
    assert ghz.depth() == 5 and ghz.depth(local=False) == 4 and ghz.size() == 5
    layers = ghz.layers()
    assert len(layers.elements) == 5 and all(len(layer.elements) == 1 for layer in layers.elements)
    two = qf.DAGCircuit([qf.CNOT(0, 1), qf.CNOT(2, 3), qf.H(0)])
    assert two.component_nb() == 2 and len(two.components()) == 2 and two.qubits == (0, 1, 2, 3)
    from quantumflow_b200 import workloads
    wb = workloads.wb_circuit(qf, 8, 5, 0)
    assert qf.DAGCircuit(wb).depth() == 1 + 2 * 5          # depth = 1 + 2D


def test_qaoa_builders(golden):
    import networkx as nx
    data = golden('qaoa.npz')
    graph = nx.from_edgelist([[0, 1], [1, 2], [1, 3]])
    assert np.array_equal(qf.graph_cuts(graph).reshape(-1), data['cuts'])
    circ = qf.qubo_circuit(graph, 5, [0.5] * 5, [0.5] * 5)
    assert circ.size() == 4 + 5 * (3 + 4)
    # element counts of tests/test_qaoa.py:30,36
    square = nx.Graph([(0, 1), (1, 2), (2, 3), (3, 0)])
    assert qf.qubo_circuit(square, 1, [1], [1]).size() == 12


def test_classification():
    assert classify.is_diagonal(qf.CZ().matrix()) and classify.is_diagonal(qf.ZZ(0.3).matrix())
    c, t, red = classify.peel_controls(qf.CNOT().matrix(), 2)
    assert (c, t) == ([0], [1]) and np.array_equal(red, qf.X().matrix())
    c, t, red = classify.peel_controls(qf.CCNOT().matrix(), 3)
    assert (c, t) == ([0, 1], [2])
    c, t, red = classify.peel_controls(qf.CSWAP().matrix(), 3)
    assert (c, t) == ([0], [1, 2]) and np.array_equal(red, qf.SWAP().matrix())
    c, t, red = classify.peel_controls(qf.CNOT(1, 0).permute([0, 1]).matrix(), 2)
    assert (c, t) == ([1], [0])
    assert classify.peel_controls(qf.SWAP().matrix(), 2)[0] == []
    kinds = {name: classify.g1_kind(qf.STDGATES[name](*PARAMS[name]).matrix())
             for name in ('X', 'Y', 'H', 'RY', 'RX', 'TX')}
    assert kinds == {'X': 3, 'Y': 4, 'H': 1, 'RY': 1, 'RX': 2, 'TX': 0}


def _emulate_partial_trace(flat, keep_bits, masks):
    """What qfb_partial_trace computes, in numpy: out[j] = sum_s in[deposit(j, keep_bits) | spread(s, masks)]."""
    nkeep, ntr = len(keep_bits), len(masks)
    j = np.arange(1 << nkeep, dtype=np.int64)
    base = np.zeros_like(j)
    for b, pos in enumerate(keep_bits):
        base |= ((j >> b) & 1) << pos
    out = np.zeros(1 << nkeep, dtype=flat.dtype)
    for s in range(1 << ntr):
        off = sum(m for t, m in enumerate(masks) if (s >> t) & 1)
        out = out + flat[base | off]
    return out


@pytest.mark.parametrize('count,rank,traced', [(3, 2, [1]), (4, 2, [0, 3]), (2, 4, [1]), (5, 2, [4, 0, 2]),
                                               (9, 3, [0, 2, 3, 5, 6, 8])])
def test_partial_trace_layout_matches_reference_einsum(count, rank, traced):
    # reference qubits.py:216-225: implicit-mode np.einsum with the traced axis letter repeated in every block;
    # (9, 3, ...) has 27 axes, so the last one is labelled 'A' and sorts FIRST in the reference's output
    from quantumflow_b200.qubits import partial_trace_layout
    from quantumflow_b200 import backend as bk
    total = count * rank
    rng = np.random.RandomState(count * 10 + rank)
    data = rng.randint(-3, 4, size=[2] * total).astype(np.int64 if total < 20 else np.int8)
    sub = list(bk.EINSUM_SUBSCRIPTS[:total])
    for ax in traced:
        for block in range(1, rank):
            sub[block * count + ax] = sub[ax]
    want = np.einsum(''.join(sub), data)
    keep_bits, masks = partial_trace_layout(count, rank, traced)
    got = _emulate_partial_trace(data.reshape(-1), keep_bits, masks).reshape(want.shape)
    assert np.array_equal(got, want)

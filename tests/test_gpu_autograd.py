"""QAOA gradient path (BASELINE.json configs[1]): expectation and gradient through the torch.autograd bridge
against the reference forward pass and its central finite differences (tests/golden/qaoa*.npz; the reference's
own autograd is TensorFlow, which is not installable -- SURVEY Appendix E)."""
import numpy as np
import pytest
import torch

import quantumflow_b200 as qf

pytestmark = pytest.mark.gpu


def _graph(edges):
    import networkx as nx
    return nx.from_edgelist([list(map(int, e)) for e in edges])


def test_qaoa_expectation_and_gradient_match_reference(golden):
    data = golden('qaoa.npz')
    graph = _graph(data['edges'])
    beta = torch.full((5,), 0.5, dtype=torch.float64, requires_grad=True)
    gamma = torch.full((5,), 0.5, dtype=torch.float64, requires_grad=True)
    circ = qf.qubo_circuit(graph, 5, beta, gamma)
    ket = circ.run()
    assert np.abs(qf.asarray(ket.tensor).reshape(-1) - data['ket']).max() < 1e-10
    expect = ket.expectation(qf.graph_cuts(graph))
    assert abs(float(expect) - data['expectation'][0]) < 1e-10           # north_star tolerance
    assert abs(float(expect) - 2.5810444679389155) < 1e-10
    expect.backward()
    # finite differences of the reference carry ~1e-9 of truncation error
    assert np.abs(beta.grad.numpy() - data['dbeta']).max() < 1e-6
    assert np.abs(gamma.grad.numpy() - data['dgamma']).max() < 1e-6
    # analytic values quoted in SURVEY Appendix E (torch-CPU autograd on the reference's einsum)
    want_b = [-0.110378587443, -0.454038438346, -0.835023221428, -1.164339136466, -0.766844081551]
    want_g = [-0.374760069504, -0.108290662705, 0.228910131850, 0.635745913682, 0.657584391366]
    assert np.abs(beta.grad.numpy() - want_b).max() < 1e-9
    assert np.abs(gamma.grad.numpy() - want_g).max() < 1e-9


def test_qaoa6_forward_matches_reference(golden):
    data = golden('qaoa6.npz')
    graph = _graph(data['edges'])
    graph.add_nodes_from(range(6))
    circ = qf.qubo_circuit(graph, 5, list(data['beta']), list(data['gamma']))
    ket = circ.run()
    assert np.abs(qf.asarray(ket.tensor).reshape(-1) - data['ket']).max() < 1e-10
    assert abs(float(qf.asarray(ket.expectation(qf.graph_cuts(graph)))) - data['expectation'][0]) < 1e-10


def test_tensormul_backward_agrees_with_torch_matmul():
    """The bridge's gradients (U^H g through the apply kernel, g psi^H through qfb_gate_grad) equal torch's own
    matmul backward on the same contraction."""
    torch.manual_seed(0)
    n = 6
    from quantumflow_b200 import backend as bk
    for indices in ([2], [5, 1], [0, 3]):
        k = len(indices)
        op = torch.randn([2] * (2 * k), dtype=torch.complex128, requires_grad=True)
        psi = torch.randn([2] * n, dtype=torch.complex128, device='cuda', requires_grad=True)
        weight = torch.randn([2] * n, dtype=torch.float64, device='cuda')
        out = bk.tensormul(op, psi, indices)
        loss = (weight * (out.abs() ** 2)).sum()
        loss.backward()
        op2 = op.detach().clone().requires_grad_(True)
        psi2 = psi.detach().cpu().clone().requires_grad_(True)
        rest = [a for a in range(n) if a not in indices]
        moved = psi2.permute(indices + rest).reshape(1 << k, -1)
        res = (op2.reshape(1 << k, 1 << k) @ moved).reshape([2] * n)
        inv = [0] * n
        for pos, ax in enumerate(indices + rest):
            inv[ax] = pos
        loss2 = (weight.cpu() * (res.permute(inv).abs() ** 2)).sum()
        loss2.backward()
        assert abs(float(loss) - float(loss2)) < 1e-10
        assert (op.grad - op2.grad).abs().max() < 1e-10
        assert (psi.grad.cpu() - psi2.grad).abs().max() < 1e-10


def test_gradient_descent_reaches_the_reference_criterion():
    """tests/test_qaoa_maxcut.py:18-19: ratio > 0.95 on the star graph [[0,1],[1,2],[1,3]] (plain GD, lr 0.01,
    initial beta, gamma ~ N(0.5, 0.01) as in examples/qaoa_maxcut.py:47-79)."""
    graph = _graph([[0, 1], [1, 2], [1, 3]])
    cuts = qf.graph_cuts(graph)
    np.random.seed(0)
    beta = torch.tensor(np.random.normal(0.5, 0.01, size=5), requires_grad=True)
    gamma = torch.tensor(np.random.normal(0.5, 0.01, size=5), requires_grad=True)
    ratio = 0.0
    for step in range(150):
        circ = qf.qubo_circuit(graph, 5, beta, gamma)
        expect = circ.run().expectation(cuts)
        ratio = float(expect) / cuts.max()
        if ratio > 0.96:
            break
        (-expect).backward()
        with torch.no_grad():
            beta -= 0.01 * beta.grad
            gamma -= 0.01 * gamma.grad
        beta.grad = None
        gamma.grad = None
    assert ratio > 0.95

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


def _random_parametrised_circuit(n, layers, seed, params):
    """RX / RY / RZ / ZZ / CPHASE gates on random qubits, every angle an entry of `params` (torch, requires_grad),
    CNOT / H without parameters in between."""
    rnd = np.random.RandomState(seed)
    circ = qf.Circuit()
    at = 0
    for _ in range(layers):
        for q in range(n):
            kind = rnd.choice(['RX', 'RY', 'RZ', 'H'])
            if kind == 'H':
                circ += qf.H(q)
            else:
                circ += getattr(qf, kind)(params[at], q)
                at += 1
        for _ in range(n // 2):
            a, b = rnd.choice(n, 2, replace=False)
            kind = rnd.choice(['ZZ', 'CNOT', 'CPHASE'])
            if kind == 'CNOT':
                circ += qf.CNOT(int(a), int(b))
            else:
                circ += getattr(qf, kind)(params[at], int(a), int(b))
                at += 1
    return circ, at


@pytest.mark.parametrize('n,layers,seed', [(3, 2, 0), (6, 4, 1), (9, 3, 2), (12, 2, 3)])
def test_single_launch_circuit_node_matches_the_gate_by_gate_path(monkeypatch, n, layers, seed):
    """autograd.run_small_circuit (one launch forward, one reverse adjoint sweep that recomputes the intermediate
    states with U^H) against the per-gate nodes (one saved state per gate): same state, same expectation, same
    gradient for every parameter and for a differentiable input state."""
    from quantumflow_b200 import engine
    rnd = np.random.RandomState(100 + seed)
    diag = rnd.standard_normal(1 << n).reshape([2] * n)
    results = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('QFB_SMALL_CIRCUIT', mode)
        params = torch.tensor(np.random.RandomState(7).uniform(0, 2 * np.pi, 400), dtype=torch.float64, requires_grad=True)
        circ, used = _random_parametrised_circuit(n, layers, seed, params)
        before = engine.launch_count()
        ket = circ.run()
        fwd_launches = engine.launch_count() - before
        expect = ket.expectation(diag)
        expect.backward()
        results[mode] = (qf.asarray(ket.tensor).reshape(-1), float(expect), params.grad.numpy()[:used].copy(),
                         fwd_launches)
    fused, plain = results['1'], results['0']
    assert np.abs(fused[0] - plain[0]).max() < 1e-12
    assert abs(fused[1] - plain[1]) < 1e-12
    assert np.abs(fused[2] - plain[2]).max() < 1e-10
    assert np.abs(plain[2]).max() > 1e-3
    assert fused[3] <= 3 < plain[3]            # zero state + ONE launch for the whole circuit


def test_small_circuit_kernels_run_a_batch_of_parameter_sets_as_one_cta_each():
    """C ABI of csrc/qfb_small.cu with batch > 1 (independent states and matrices per item, shared gate list): forward
    against the oracle's gate-by-gate arithmetic, adjoint against the per-item results of the batch-1 call."""
    import ctypes
    from oracle import qf_oracle as O
    from quantumflow_b200 import _lib, engine
    lib = _lib.load()
    n, batch = 5, 3
    rnd = np.random.RandomState(5)
    names = [('RX', 1, (0,)), ('ZZ', 1, (1, 3)), ('H', 0, (2,)), ('CNOT', 0, (4, 0)), ('RY', 1, (3,)), ('CPHASE', 1, (2, 4))]
    rows, at = [], 0
    for name, _, qubits in names:
        bits = [n - 1 - q for q in qubits]
        rows.append([len(bits), bits[0], bits[1] if len(bits) > 1 else 0, at])
        at += 1 << (2 * len(bits))
    stride = at
    mats = np.zeros((batch, stride), dtype=np.complex128)
    states = np.zeros((batch, 1 << n), dtype=np.complex128)
    want = np.zeros_like(states)
    for b in range(batch):
        psi = rnd.standard_normal(1 << n) + 1j * rnd.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        states[b] = psi
        cur = psi.copy()
        for (name, npar, qubits), row in zip(names, rows):
            m = O.gate_matrix(name, tuple(rnd.uniform(0, 6, npar)))
            mats[b, row[3]:row[3] + m.size] = m.reshape(-1)
            cur = O.tensormul_flat(m, cur, [n - 1 - q for q in qubits])
        want[b] = cur
    dev = torch.device('cuda')
    d_states = torch.from_numpy(states).to(dev)
    d_mats = torch.from_numpy(mats).to(dev)
    d_desc = torch.tensor(rows, dtype=torch.int32, device=dev)
    out = torch.empty_like(d_states)
    st = engine._stream()
    _lib.check(lib.qfb_small_circuit_run(out.data_ptr(), d_states.data_ptr(), n, batch, len(rows), d_desc.data_ptr(),
                                         d_mats.data_ptr(), stride, st))
    assert np.abs(out.cpu().numpy() - want).max() < 1e-13
    grad_out = torch.from_numpy(rnd.standard_normal((batch, 1 << n)) + 1j * rnd.standard_normal((batch, 1 << n))).to(dev)

    def adjoint(psi_final, g, m, nb):
        gm = torch.zeros((nb, stride), dtype=torch.complex128, device=dev)
        gin = torch.empty_like(psi_final)
        scratch = torch.empty(int(lib.qfb_small_circuit_scratch_doubles(nb, len(rows))), dtype=torch.float64, device=dev)
        _lib.check(lib.qfb_small_circuit_adjoint(psi_final.data_ptr(), g.data_ptr(), n, nb, len(rows), d_desc.data_ptr(),
                                                 m.data_ptr(), stride, gm.data_ptr(), gin.data_ptr(),
                                                 scratch.data_ptr(), st))
        return gm.cpu().numpy(), gin.cpu().numpy()

    gm_all, gin_all = adjoint(out, grad_out, d_mats, batch)
    for b in range(batch):
        gm_one, gin_one = adjoint(out[b:b + 1].contiguous(), grad_out[b:b + 1].contiguous(), d_mats[b:b + 1].contiguous(), 1)
        assert np.array_equal(gm_all[b], gm_one[0]) and np.array_equal(gin_all[b], gin_one[0])
        # grad_in = U_1^H ... U_G^H grad_out: check against the oracle applied in reverse
        cur = grad_out[b].cpu().numpy()
        for (name, npar, qubits), row in reversed(list(zip(names, rows))):
            dim = 1 << row[0]
            m = mats[b, row[3]:row[3] + dim * dim].reshape(dim, dim)
            cur = O.tensormul_flat(m.conj().T.copy(), cur, [n - 1 - q for q in qubits])
        assert np.abs(gin_all[b] - cur).max() < 1e-12

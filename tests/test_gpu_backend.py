"""GPU parity: the backend contract (tensormul / inner / outer / productdiag ...) and the C-ABI kernels
against the reference-generated fixtures and the oracle. Bar: 1e-10 max-abs (complex128)."""
import itertools
import json
import os

import numpy as np
import pytest
import torch

from oracle import qf_oracle as O

from conftest import AMP_TOL, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def bk():
    from quantumflow_b200 import backend
    return backend


@pytest.fixture(scope='module')
def engine():
    from quantumflow_b200 import engine
    return engine


def dev(array):
    return torch.from_numpy(np.ascontiguousarray(array, dtype=np.complex128)).cuda()


def host(tensor):
    return tensor.detach().cpu().numpy()


def rand_state(rng, n):
    return rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)


def test_tensormul_golden_cases(bk, golden):
    """bk.tensormul with K = 1..6, unsorted indices, against the reference's own outputs."""
    data = golden('tensormul.npz')
    cases = json.load(open(os.path.join(GOLDEN, 'tensormul_cases.json')))
    for case in cases:
        gate, state, want = (data[case['tag'] + s] for s in ('_gate', '_state', '_out'))
        got = bk.tensormul(bk.astensor(gate), bk.asamplitudes(state), case['indices'])
        assert got.is_cuda and tuple(got.shape) == want.shape
        assert np.abs(host(got) - want).max() < AMP_TOL, case


def test_inner_outer_productdiag_trace_golden(bk, golden):
    data = golden('tensormul.npz')
    a, b = bk.asamplitudes(data['inner_a']), bk.asamplitudes(data['inner_b'])
    assert abs(complex(host(bk.inner(a, b))) - complex(data['inner_out'])) < AMP_TOL
    got = bk.outer(bk.asamplitudes(data['inner_a'][0, 0]), bk.asamplitudes(data['inner_b'][1, 1, 0]))
    assert np.abs(host(got) - data['outer_out']).max() < AMP_TOL
    rho = bk.asamplitudes(data['productdiag_in'])
    assert np.abs(host(bk.productdiag(rho)) - data['productdiag_out']).max() == 0.0
    mat = data['productdiag_in'].reshape(8, 8)
    assert abs(complex(host(bk.trace(bk.reshape(rho, [8, 8])))) - np.trace(mat)) < AMP_TOL
    # transpose / conj on amplitude tensors are kernels too
    perm = [2, 0, 5, 1, 4, 3]
    assert np.array_equal(host(bk.transpose(rho, perm)), np.transpose(data['productdiag_in'], perm))
    assert np.array_equal(host(bk.conj(rho)), np.conj(data['productdiag_in']))
    # backend tests of the reference (tests/test_backend.py:65-87): inner vs np.vdot, outer vs np.outer
    rng = np.random.RandomState(0)
    for n in (1, 3, 8):
        x, y = rand_state(rng, n), rand_state(rng, n)
        assert abs(complex(host(bk.inner(dev(x), dev(y)))) - np.vdot(x, y)) < AMP_TOL


@pytest.mark.parametrize('n', [1, 2, 5, 11])
def test_every_bit_position_dense_1q_2q(engine, n):
    rng = np.random.RandomState(n)
    psi = rand_state(rng, n)
    m1 = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    for b in range(n):
        got = host(engine.apply_operator(dev(psi), m1, [b]))
        assert np.abs(got - O.tensormul_flat(m1, psi, [b])).max() < AMP_TOL
    m2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    for b0, b1 in itertools.permutations(range(n), 2):
        if n == 11 and (b0 + 3 * b1) % 7:
            continue
        got = host(engine.apply_operator(dev(psi), m2, [b0, b1]))
        assert np.abs(got - O.tensormul_flat(m2, psi, [b0, b1])).max() < AMP_TOL


def test_inplace_controls_diag_and_generic_paths(engine):
    rng = np.random.RandomState(7)
    n = 10
    psi = rand_state(rng, n)
    for name, bits in [('CNOT', [3, 8]), ('CNOT', [0, 9]), ('CZ', [9, 0]), ('CCNOT', [2, 7, 4]), ('CSWAP', [9, 1, 5]),
                       ('SWAP', [0, 1]), ('ISWAP', [6, 2]), ('T', [0]), ('RZ', [9]), ('ZZ', [4, 5])]:
        params = (0.37,) if name in ('RZ', 'ZZ') else ()
        mat = O.gate_matrix(name, params)
        want = O.tensormul_flat(mat, psi, bits)
        out = engine.apply_operator(dev(psi), mat, bits)
        assert np.abs(host(out) - want).max() < AMP_TOL, name
        buf = dev(psi)
        same = engine.apply_operator(buf, mat, bits, inplace=True)
        assert same.data_ptr() == buf.data_ptr() and np.abs(host(buf) - want).max() < AMP_TOL, name
        raw = engine.apply_operator(dev(psi), mat, bits, classify_structure=False)    # dense path, no peeling
        assert np.abs(host(raw) - want).max() < AMP_TOL, name
    # generic kernel: k = 5..8, in and out of place
    for k in (5, 6, 7, 8):
        bits = list(rng.permutation(n)[:k])
        mat = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        want = O.tensormul_flat(mat, psi, bits)
        assert np.abs(host(engine.apply_operator(dev(psi), mat, bits)) - want).max() < 1e-9
        buf = dev(psi)
        engine.apply_operator(buf, mat, bits, inplace=True)
        assert np.abs(host(buf) - want).max() < 1e-9
    # whole-state operator (k == n) and scalar (k == 0)
    small = rand_state(rng, 3)
    m = rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8))
    assert np.abs(host(engine.apply_operator(dev(small), m, [2, 1, 0])) - m @ small).max() < AMP_TOL
    # errors: library reports, python raises
    from quantumflow_b200._lib import QfbError
    with pytest.raises(QfbError):
        engine.apply_operator(dev(psi), O.gate_matrix('H'), [n])
    with pytest.raises(QfbError):
        engine.apply_operator(dev(psi), m2 := np.ones((4, 4)), [3, 3])


def test_index_hi_resolution_for_sharded_states(engine):
    """Diagonal entries / controls that live in the rank bits (positions >= nbits)."""
    rng = np.random.RandomState(9)
    n, p = 9, 2
    nl = n - p
    full = rand_state(rng, n)
    for name, bits in [('CZ', [n - 1, 3]), ('CNOT', [n - 2, 0]), ('ZZ', [n - 1, n - 2]), ('T', [n - 1]),
                       ('CCNOT', [n - 1, 2, 5]), ('CPHASE', [4, n - 2])]:
        params = (0.41,) if name in ('ZZ', 'CPHASE') else ()
        mat = O.gate_matrix(name, params)
        want = O.tensormul_flat(mat, full, bits)
        got = np.empty_like(full)
        for rank in range(1 << p):
            shard = dev(full[rank << nl:(rank + 1) << nl])
            got[rank << nl:(rank + 1) << nl] = host(engine.apply_operator(shard, mat, bits, index_hi=rank))
        assert np.abs(got - want).max() < AMP_TOL, name


def test_reductions_and_elementwise(engine):
    rng = np.random.RandomState(3)
    for n in (0, 1, 4, 13, 18):
        psi = rand_state(rng, n)
        d = dev(psi)
        assert abs(float(engine.norm2(d)) - np.vdot(psi, psi).real) < 1e-9 * max(1, psi.size)
        other = rand_state(rng, n)
        assert abs(complex(host(engine.vdot(d, dev(other)))) - np.vdot(psi, other)) < 1e-9 * max(1, psi.size)
        assert np.abs(host(engine.probabilities(d)) - np.abs(psi) ** 2).max() < 1e-13
        diag = rng.normal(size=psi.size)
        want = np.sum(diag * np.abs(psi) ** 2)
        assert abs(float(engine.expectation_diag(d, torch.from_numpy(diag).cuda())) - want) < 1e-9 * max(1, psi.size)
        nrm = host(engine.normalize_by_norm2(d))
        assert np.abs(nrm - psi / np.sqrt(np.vdot(psi, psi).real)).max() < 1e-13
        assert np.abs(host(engine.scale(d, 0.3 - 2j)) - (0.3 - 2j) * psi).max() < 1e-13
        assert np.abs(host(engine.axpby(d, 2.0, dev(other), -1j)) - (2 * psi - 1j * other)).max() < 1e-13
        assert np.abs(host(engine.axpby(d, 1j)) - 1j * psi).max() < 1e-13
        for bit in range(min(n, 3)):
            mask = ((np.arange(psi.size) >> bit) & 1).astype(bool)
            p = np.abs(psi) ** 2
            got = host(engine.marginal(d, bit))
            assert abs(got[0] - p[~mask].sum()) < 1e-9 * psi.size and abs(got[1] - p[mask].sum()) < 1e-9 * psi.size
            col = host(engine.collapse(d, bit, 1, 0.5))
            assert np.abs(col - np.where(mask, 0.5 * psi, 0)).max() < 1e-14
    # outer with conjugation (State.asdensity) and density read-out
    a, b = rand_state(rng, 5), rand_state(rng, 5)
    rho = host(engine.outer(dev(a), dev(b), conj_second=True)).reshape(32, 32)
    assert np.abs(rho - np.outer(a, b.conj())).max() < 1e-14
    assert abs(complex(host(engine.density_trace(dev(rho), 5))) - np.trace(rho)) < 1e-12
    assert np.array_equal(host(engine.density_diag(dev(rho), 5)).reshape(-1), np.diag(rho))
    scalar = torch.tensor(2.0 - 1.0j, dtype=torch.complex128, device='cuda')
    assert np.abs(host(engine.divide_by_device_scalar(dev(a), scalar)) - a / (2 - 1j)).max() < 1e-14
    # bit permutation: dst bit j <- src bit perm[j]
    n = 9
    psi = rand_state(rng, n)
    perm = list(rng.permutation(n))
    got = host(engine.permute_bits(dev(psi), perm, conj=True))
    idx = np.arange(1 << n)
    src = np.zeros_like(idx)
    for j in range(n):
        src |= ((idx >> j) & 1) << perm[j]
    assert np.array_equal(got, np.conj(psi[src]))


def test_sample_search_matches_sequential_cdf(engine):
    rng = np.random.RandomState(5)
    for n in (3, 10, 15):
        p = rng.random(1 << n)
        p /= p.sum()
        u = rng.random(64)
        got = engine.sample_search(torch.from_numpy(p).cuda(), u)
        cdf = np.cumsum(p)
        want = np.searchsorted(cdf / cdf[-1], u, side='right')
        # block-hierarchical sums vs one sequential cumsum: identical except when u sits within rounding of an edge
        assert np.mean(got == want) > 0.95 and np.abs(got.astype(np.int64) - want).max() <= 1


def test_gate_grad_kernel(engine):
    rng = np.random.RandomState(2)
    n = 8
    g, psi = rand_state(rng, n), rand_state(rng, n)
    for bits in ([3], [0], [7, 2], [1, 6], [5, 0, 3]):
        k = len(bits)
        got = host(engine.gate_grad(dev(g), dev(psi), bits))
        axes = [n - 1 - b for b in bits]
        gt = np.moveaxis(g.reshape([2] * n), axes, range(k)).reshape(1 << k, -1)
        pt = np.moveaxis(psi.reshape([2] * n), axes, range(k)).reshape(1 << k, -1)
        assert np.abs(got - gt @ pt.conj().T).max() < 1e-10


def test_no_cpu_path_for_amplitudes(bk):
    """Amplitude tensors only exist on the device; operators stay on the host."""
    import quantumflow_b200 as qf
    ket = qf.zero_state(3)
    assert ket.tensor.is_cuda and not qf.H(0).tensor.is_cuda
    assert qf.H(0).run(ket).tensor.is_cuda
    assert bk.gpu_available() and bk.DEVICE == 'gpu'


def test_plugin_backend_for_the_reference_classes_runs_on_the_device(bk):
    """b200ref is the module the reference's backend package star-imports: everything the
    reference builds through bk.astensorproduct (quantumflow/qubits.py:103) is an amplitude tensor in HBM and every
    tensormul launches a libqfb200 kernel -- no silent CPU path (round-1 review, missing #1)."""
    from oracle import qf_oracle as O
    from quantumflow_b200 import engine
    from quantumflow_b200.backend import b200ref as ref
    rng = np.random.RandomState(3)
    n = 7
    psi = rng.normal(size=[2] * n) + 1j * rng.normal(size=[2] * n)
    ket = ref.astensorproduct(psi)                 # what quantumflow.states.State.__init__ calls
    assert ket.is_cuda and tuple(ket.shape) == (2,) * n
    gate = ref.astensorproduct(O.gate_matrix('CNOT', ()))     # what quantumflow.ops.Gate.__init__ calls
    assert gate.is_cuda and tuple(gate.shape) == (2,) * 4
    before = engine.launch_count()
    out = ref.tensormul(gate, ket, [5, 2])
    assert engine.launch_count() > before and out.is_cuda
    want = O.tensormul(O.as_tensor(O.gate_matrix('CNOT', ())), psi, [5, 2])
    assert np.abs(ref.evaluate(out) - want).max() < 1e-12
    # a gate tensor derived on the device (no remembered host copy) and a host tensor1 both end up in the kernels
    rx = ref.astensorproduct(O.gate_matrix('RX', (0.3,)))
    before = engine.launch_count()
    out2 = ref.tensormul(ref.conj(rx), psi, [4])
    assert engine.launch_count() > before and out2.is_cuda
    want2 = O.tensormul(O.as_tensor(O.gate_matrix('RX', (0.3,)).conj()), psi, [4])
    assert np.abs(ref.evaluate(out2) - want2).max() < 1e-12
    # gate (x) gate as the reference's Gate.__matmul__ does it (ops.py:187-198): tensor1 is a [2]*2K gate tensor
    cz = ref.astensorproduct(O.gate_matrix('CZ', ()))
    prod = ref.tensormul(rx, cz, [1])
    want3 = O.tensormul(O.as_tensor(O.gate_matrix('RX', (0.3,))), O.as_tensor(O.gate_matrix('CZ', ())), [1])
    assert np.abs(ref.evaluate(prod) - want3).max() < 1e-12
    assert abs(complex(ref.evaluate(ref.inner(psi, psi))) - np.vdot(psi, psi)) < 1e-10
    # the mirror package's backend refuses a state-sized tensor on the host instead of multiplying it on the CPU
    big = np.zeros([2] * 17, dtype=np.complex128)
    with pytest.raises(RuntimeError):
        bk.tensormul(bk.astensorproduct(O.gate_matrix('H', ())), bk.astensorproduct(big), [3])


def test_staged_transfer_of_large_host_arrays(bk):
    """State(host array) / asarray(result) of state-sized arrays go through the pinned staging ring (engine.upload /
    engine.download): byte-exact round trip, several pieces, ragged last piece."""
    from quantumflow_b200 import engine
    rng = np.random.RandomState(2)
    old = engine._STAGE_BYTES
    try:
        engine._STAGE_BYTES = 1 << 20                  # 64 Ki amplitudes per piece: many pieces at test size
        engine._stage_buffers.clear()
        n = (1 << 18) + 12345                          # not a multiple of the piece
        host = rng.normal(size=n) + 1j * rng.normal(size=n)
        dev = engine.upload(host, bk.device())
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)
        assert np.array_equal(engine.download(dev), host)
    finally:
        engine._STAGE_BYTES = old
        engine._stage_buffers.clear()
    # the public path: a 2^24-amplitude numpy array (256 MiB) becomes a resident State and comes back unchanged
    import quantumflow_b200 as qf
    big = rng.normal(size=1 << 24) + 1j * rng.normal(size=1 << 24)
    ket = qf.State(big.reshape([2] * 24))
    assert ket.tensor.is_cuda
    assert np.array_equal(qf.asarray(ket.tensor).reshape(-1), big)

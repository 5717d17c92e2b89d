"""GPU parity of the State / Density / Gate / Channel / Measure API against reference-generated fixtures.
The test bodies follow the reference's own tests (cited) so that they read like its conformance suite."""
import numpy as np
import pytest

import quantumflow_b200 as qf
from quantumflow_b200 import workloads
from oracle import qf_oracle as O

from conftest import AMP_TOL

pytestmark = pytest.mark.gpu


def amps(state):
    return qf.asarray(state.tensor).reshape(-1)


def test_standard_states():
    # reference tests/test_states.py:25-41
    ket = qf.zero_state(4)
    assert ket.vec.asarray()[0, 0, 0, 0] == 1 and qf.asarray(ket.norm()) == 1
    w = qf.w_state(4).vec.asarray()
    assert w[1, 0, 0, 0] == w[0, 1, 0, 0] == w[0, 0, 0, 1] == 0.5 and w[0, 0, 0, 0] == 0
    ghz = qf.ghz_state(3).vec.asarray()
    assert abs(ghz[0, 0, 0] - 1 / np.sqrt(2)) < 1e-15 and abs(ghz[1, 1, 1] - 1 / np.sqrt(2)) < 1e-15
    assert qf.zero_state(['a', 'b']).qubits == ('a', 'b')
    with pytest.raises(ValueError):
        qf.State(np.zeros(8), qubits=[0, 1])
    assert abs(float(qf.asarray(qf.random_state(5).norm())) - 1) < 1e-14
    probs = qf.asarray(qf.random_state(3).probabilities())
    assert probs.shape == (2, 2, 2) and abs(probs.sum() - 1) < 1e-14


def test_gate_bits_and_immutability():
    # reference tests/test_gates.py:59-77: X on qubit i sets axis i
    for n in (1, 3, 6):
        for i in range(n):
            ket0 = qf.zero_state(n)
            ket = qf.X(i).run(ket0)
            idx = [0] * n
            idx[i] = 1
            assert ket.vec.asarray()[tuple(idx)] == 1
            assert ket0.vec.asarray()[(0,) * n] == 1           # input untouched
    ket = qf.zero_state([7, 'q', 3])
    ket = qf.X('q').run(ket)
    assert ket.vec.asarray()[0, 1, 0] == 1
    with pytest.raises(ValueError):
        qf.X(5).run(qf.zero_state(2))
    with pytest.raises(TypeError):
        qf.X(0).aschannel().run(qf.zero_state(1))


def test_gate_run_matches_oracle_for_every_standard_gate(golden):
    import json
    import os
    from conftest import GOLDEN
    params = json.load(open(os.path.join(GOLDEN, 'stdgates_params.json')))
    np.random.seed(1)
    n = 5
    ket = qf.random_state(n)
    psi = amps(ket)
    for name, p in params.items():
        cls = qf.STDGATES[name]
        k = cls(*p).qubit_nb
        qubits = list(np.random.permutation(n)[:k])
        got = amps(cls(*p, *qubits).run(ket))
        want = O.tensormul(O.as_tensor(O.gate_matrix(name, p)), psi.reshape([2] * n), [int(q) for q in qubits])
        assert np.abs(got - want.reshape(-1)).max() < AMP_TOL, name


def test_expectation_and_projectors():
    # reference tests/test_gates.py:150-164 (42 and 2.5) and :116-126
    ket = qf.zero_state(1)
    ket = qf.H(0).run(ket)
    m = ket.expectation([0.4, 0.6])
    assert abs(float(qf.asarray(m)) - 0.5) < 1e-14
    ket = qf.zero_state(2)
    for gate in (qf.H(0), qf.CNOT(0, 1)):
        ket = gate.run(ket)
    assert abs(float(qf.asarray(ket.expectation([[42, 0], [0, 42]]))) - 42) < 1e-12
    assert abs(float(qf.asarray(ket.expectation([[2, 0], [0, 3]]))) - 2.5) < 1e-14
    half = qf.P0(0).run(qf.H(0).run(qf.zero_state(1)))
    assert abs(float(qf.asarray(half.norm())) - 0.5) < 1e-15
    np.random.seed(3)
    est = float(qf.asarray(ket.expectation([[2, 0], [0, 3]], trials=2000)))
    assert abs(est - 2.5) < 0.1


def test_sampling_is_bit_exact_under_the_shared_rng_stream(golden):
    """Same seed, same numpy global-RNG calls => same outcomes as the reference (states.py:121-160)."""
    data = golden('sampling.npz')
    ket = workloads.wb_circuit(qf, 8, 3, 9).run()
    np.random.seed(42)
    seq = np.asarray([ket.measure() for _ in range(16)])
    assert np.array_equal(seq, data['measure_seq'])
    np.random.seed(43)
    assert np.array_equal(ket.sample(1000).reshape(-1), data['sample_1000'])


def test_midcircuit_measurement_matches_reference(golden):
    data = golden('sampling.npz')
    np.random.seed(44)
    ro = qf.Register('ro')
    prog = qf.Circuit([qf.H(q) for q in range(6)])
    prog += qf.CNOT(0, 1)
    prog += qf.Measure(0, ro[0])
    prog += qf.RX(0.3, 2)
    prog += qf.CNOT(2, 3)
    prog += qf.Measure(3, ro[1])
    prog += qf.Measure(1, ro[2])
    res = prog.run()
    assert [res.memory[ro[i]] for i in range(3)] == list(data['midcircuit_bits'])
    assert np.abs(amps(res) - data['midcircuit_state']).max() < AMP_TOL
    assert np.random.random() == data['midcircuit_next_random'][0]      # exactly three draws were consumed
    assert res.cbits == (ro[0], ro[1], ro[2]) and res.cbit_nb == 3


def test_stochastic_unravelling_and_density_measurement(golden):
    data = golden('sampling.npz')
    np.random.seed(45)
    ket = qf.Circuit([qf.H(0), qf.CNOT(0, 1), qf.RY(0.4, 2)]).run(qf.zero_state(3))
    for q in range(3):
        ket = qf.Damping(0.3, q).run(ket)
        ket = qf.Depolarizing(0.5, q).run(ket)
    assert np.abs(amps(ket) - data['kraus_run_state']).max() < AMP_TOL
    np.random.seed(46)
    ro = qf.Register('ro')
    rho = qf.Circuit([qf.H(0), qf.CNOT(0, 1), qf.RX(0.7, 1)]).evolve()
    rho = qf.Measure(1, ro[0]).evolve(rho)
    assert np.abs(qf.asarray(rho.asoperator()) - data['measure_evolve_rho']).max() < AMP_TOL
    assert rho.memory[ro[0]] == data['measure_evolve_bit'][0]
    np.random.seed(47)
    assert np.abs(amps(qf.random_state(4)) - data['random_state4']).max() < 1e-14
    assert np.abs(qf.asarray(qf.random_density(2).asoperator()) - data['random_density2']).max() < 1e-14
    assert np.abs(amps(qf.Reset(1, 3).run(qf.State(data['reset_in']))) - data['reset_out']).max() < AMP_TOL


def test_density_api():
    # reference tests/test_states.py:143-241
    ket = qf.random_state(3)
    rho = ket.asdensity()
    assert rho.tensor.is_cuda and tuple(rho.tensor.shape) == (2,) * 6
    assert abs(complex(qf.asarray(rho.trace())) - 1) < 1e-14
    assert np.abs(qf.asarray(rho.asoperator()) - np.outer(amps(ket), amps(ket).conj())).max() < 1e-15
    assert np.abs(np.real(qf.asarray(rho.probabilities())) - qf.asarray(ket.probabilities())).max() < 1e-15
    assert abs(float(np.real(qf.asarray(qf.purity(rho)))) - 1) < 1e-13
    mixed = qf.mixed_density(2)
    assert abs(float(np.real(qf.asarray(qf.purity(mixed)))) - 0.25) < 1e-15
    assert qf.densities_close(qf.Density(qf.asarray(rho.asoperator()) * 2).normalize(), rho)
    # partial trace (read-out path)
    bell = qf.Circuit([qf.H(0), qf.CNOT(0, 1)]).run(qf.zero_state(3)).asdensity()
    red = bell.partial_trace([1, 2])
    assert red.qubits == (0,) and np.allclose(qf.asarray(red.asoperator()), np.eye(2) / 2)
    joined = qf.join_densities(qf.mixed_density([0]), qf.zero_state([1]).asdensity())
    assert joined.qubits == (0, 1) and np.allclose(np.diag(qf.asarray(joined.asoperator())), [0.5, 0, 0.5, 0])
    js = qf.join_states(qf.zero_state([0]), qf.X(1).run(qf.zero_state([1])))
    assert js.vec.asarray()[0, 1] == 1
    perm = qf.X(0).run(qf.zero_state(3)).permute([2, 1, 0])
    assert perm.vec.asarray()[0, 0, 1] == 1 and perm.qubits == (2, 1, 0)
    assert qf.states_close(ket, ket.relabel([0, 1, 2])) and not qf.states_close(ket, qf.zero_state(3))


def test_gate_and_channel_evolve_against_oracle():
    np.random.seed(5)
    n = 4
    rho = qf.random_density(n)
    r = qf.asarray(rho.tensor)
    cases = [qf.RX(0.3, 2), qf.H(0), qf.T(3), qf.CNOT(3, 1), qf.CZ(0, 2), qf.CAN(0.1, 0.2, 0.3, 2, 0),
             qf.CCNOT(1, 3, 0), qf.SWAP(0, 3)]
    for gate in cases:
        u = gate.matrix()
        k = gate.qubit_nb
        idx = [rho.qubits.index(q) for q in gate.qubits]
        want = O.tensormul(np.kron(u, u.conj()).reshape([2] * (4 * k)), r, idx + [i + n for i in idx])
        got = qf.asarray(gate.evolve(rho).tensor)
        assert np.abs(got - want).max() < AMP_TOL, gate.name
        got2 = qf.asarray(gate.aschannel().evolve(rho).tensor)
        assert np.abs(got2 - want).max() < AMP_TOL, gate.name
    for kraus in (qf.Depolarizing(0.2, 1), qf.Damping(0.1, 3), qf.Dephasing(0.4, 0)):
        got = qf.asarray(kraus.evolve(rho).tensor)
        q = kraus.qubits[0]
        want = O.tensormul(kraus.superoperator_matrix().reshape([2] * 4), r, [q, q + n])
        assert np.abs(got - want).max() < AMP_TOL
        assert qf.densities_close(kraus.aschannel().evolve(rho), kraus.evolve(rho))
        assert kraus.evolve(rho.update({qf.Register()[0]: 1})).memory == {}     # channels.py:85 drops memory


def test_reference_channel_golden_values():
    # amplitude damping, reference tests/test_channels.py:203-215
    rho = qf.zero_state(1).asdensity()
    p = 1.0 - np.exp(-50 / 15000)
    chan = qf.Damping(p, 0).aschannel()
    rho1 = chan.evolve(rho)
    assert qf.densities_close(rho, rho1)
    rho3 = chan.evolve(qf.X(0).aschannel().evolve(rho1))
    assert qf.densities_close(qf.Density([[0.00332778, 0], [0, 0.99667222]]), rho3)
    # depolarizing, reference tests/test_channels.py:218-242
    p = 1.0 - np.exp(-1 / 20)
    chan = qf.Depolarizing(p, 0).aschannel()
    rho2 = qf.Density([[0.43328691, 0.48979689], [0.48979689, 0.56671309]])
    want = qf.Density([[0.43762509, 0.45794666], [0.45794666, 0.56237491]])
    assert qf.densities_close(chan.evolve(rho2), want)
    # biased coin, reference tests/test_channels.py:159-165
    rho = qf.Circuit([qf.RX(np.pi / 3, 0)]).evolve()
    probs = np.real(qf.asarray(rho.probabilities()))
    assert np.allclose(probs, [0.75, 0.25])
    # stochastic Kraus.run vs channel, reference tests/test_channels.py:267-287 (statistical, tol 0.05)
    np.random.seed(8)
    ket0 = qf.Circuit([qf.RY(1.1, 0)]).run(qf.zero_state(1))
    kraus = qf.Damping(0.4, 0)
    acc = np.zeros((2, 2), dtype=complex)
    reps = 400
    for _ in range(reps):
        out = amps(kraus.run(ket0))
        acc += np.outer(out, out.conj()) / reps
    exact = qf.asarray(kraus.evolve(ket0.asdensity()).asoperator())
    assert np.abs(acc - exact).max() < 0.07


@pytest.mark.parametrize('n,traced', [(3, [1, 2]), (4, [2]), (8, [5]), (8, [6, 0]), (8, [0, 1, 2, 3, 4, 6, 7]),
                                      (9, [3, 8, 1])])
def test_partial_trace_on_device_matches_reference_einsum(n, traced):
    # reference qubits.py:201-227 (np.einsum with repeated subscripts on the host copy); here qfb_partial_trace:
    # one CTA per output element for small results, one thread per output element from 4096 outputs on
    np.random.seed(n)
    rho = qf.random_density(n)
    host = qf.asarray(rho.tensor)
    letters = list('abcdefghijklmnopqrstuvwxyz'[:2 * n])
    for q in traced:
        letters[n + q] = letters[q]
    want = np.einsum(''.join(letters), host)
    before = qf.engine.launch_count()
    red = rho.partial_trace(traced)
    assert red.tensor.is_cuda and red.qubits == tuple(q for q in range(n) if q not in traced)
    got = qf.asarray(red.tensor)
    assert got.shape == want.shape and np.abs(got - want).max() < AMP_TOL
    assert abs(complex(qf.asarray(red.trace())) - 1) < 1e-12
    assert qf.engine.launch_count() > before


def test_spectral_density_measures():
    # reference tests/test_measures.py:59-84, 131-156
    np.random.seed(11)
    rho0, rho1 = qf.random_density(4), qf.random_density(4)
    assert 0.0 <= qf.fidelity(rho0, rho1) <= 1.0
    assert 0.0 <= qf.fidelity(rho0, qf.random_density([3, 2, 1, 0])) <= 1.0
    assert abs(qf.fidelity(rho0, rho0) - 1) < 1e-9
    ket0, ket1 = qf.random_state(3), qf.random_state(3)
    want = float(qf.asarray(qf.state_fidelity(ket0, ket1)))
    assert abs(qf.fidelity(ket0.asdensity(), ket1.asdensity()) - want) < 1e-7
    # against the reference's own formula (scipy sqrtm), measures.py:86
    from scipy.linalg import sqrtm
    op0, op1 = qf.asarray(rho0.asoperator()), qf.asarray(rho1.asoperator())
    ref = np.real(np.trace(sqrtm(sqrtm(op0) @ op1 @ sqrtm(op0))) ** 2)
    assert abs(qf.fidelity(rho0, rho1) - ref) < 1e-9
    assert abs(qf.bures_angle(rho0, rho1) - np.arccos(np.sqrt(ref))) < 1e-8
    assert abs(qf.bures_distance(rho0, rho1) - np.sqrt(2 - 2 * np.sqrt(ref))) < 1e-8
    mixed = qf.mixed_density(4)
    assert np.isclose(qf.entropy(mixed, base=2), 4)
    assert np.isclose(qf.entropy(qf.random_gate(4).aschannel().evolve(mixed), base=2), 4)
    assert np.isclose(qf.entropy(ket0.asdensity()), 0, atol=1e-9)
    info0 = qf.mutual_info(mixed, qubits0=[0, 1], qubits1=[2, 3])
    local = qf.random_gate(2).aschannel().evolve(mixed)
    assert np.isclose(info0, qf.mutual_info(local, qubits0=[0, 1], qubits1=[2, 3]))
    assert np.isclose(info0, qf.mutual_info(local, qubits0=[0, 1]))
    bell = qf.Circuit([qf.H(0), qf.CNOT(0, 1)]).run(qf.zero_state(2)).asdensity()
    assert np.isclose(qf.mutual_info(bell, [0], [1], base=2), 2.0)


def test_state_dump_round_trip(tmp_path):
    # chunked on-disk dump of device-resident states (host-side layout tests: test_stateio_host.py)
    from quantumflow_b200 import stateio
    np.random.seed(21)
    ket = qf.random_state(12)
    path = str(tmp_path / 'ket.qfb')
    stateio.save_state(ket, path, chunk_bytes=16 * 1000)            # 5 chunks through the pinned staging buffer
    assert stateio.read_header(path) == (12, 1, 1 << 12)
    assert np.array_equal(np.fromfile(path, dtype=np.complex128, offset=stateio.HEADER_BYTES),
                          qf.asarray(ket.tensor).reshape(-1))
    back = stateio.load_state(path, chunk_bytes=16 * 700)
    assert back.tensor.is_cuda and back.qubits == ket.qubits
    assert np.array_equal(qf.asarray(back.tensor), qf.asarray(ket.tensor))
    rho = qf.random_density(4)
    path = str(tmp_path / 'rho.qfb')
    stateio.save_state(rho, path)
    back = stateio.load_state(path, qubits=rho.qubits)
    assert isinstance(back, qf.Density) and back.qubit_nb == 4
    assert np.array_equal(qf.asarray(back.tensor), qf.asarray(rho.tensor))


def test_pyquil_wavefunction_order():
    """quantumflow/forest/__init__.py:350-358: the pyQuil vector is the [2]*N tensor with reversed axes, flattened."""
    from quantumflow_b200 import stateio
    rng = np.random.RandomState(4)
    n = 9
    psi = rng.normal(size=[2] * n) + 1j * rng.normal(size=[2] * n)
    ket = qf.State(psi)
    got = stateio.state_to_wavefunction_amplitudes(ket)
    want = psi.transpose().reshape(psi.size)           # the reference's two lines
    assert np.array_equal(got, want)
    back = stateio.state_from_pyquil_order(got)
    assert np.array_equal(qf.asarray(back.tensor), psi)
    # |q0 q1 q2> = |1 0 0> : QuantumFlow index 4, pyQuil index 1
    ket = qf.Circuit([qf.X(0)]).run(qf.zero_state(3))
    assert np.argmax(np.abs(stateio.state_to_wavefunction_amplitudes(ket))) == 1


def test_purity_is_trace_of_rho_squared_also_for_non_hermitian_tensors():
    """quantumflow/measures.py:59-64 computes tr(rho . rho) as a complex scalar; sum |rho_ij|^2 only equals it for
    Hermitian rho (review finding): checked on a non-Hermitian tensor."""
    rng = np.random.RandomState(6)
    n = 4
    mat = rng.normal(size=(1 << n, 1 << n)) + 1j * rng.normal(size=(1 << n, 1 << n))
    rho = qf.Density(mat.reshape([2] * (2 * n)))
    got = complex(qf.asarray(qf.purity(rho)))
    assert abs(got - np.trace(mat @ mat)) < 1e-10
    assert abs(got - np.sum(np.abs(mat) ** 2)) > 1.0

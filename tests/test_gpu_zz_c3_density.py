"""Parity of BASELINE.json's density configuration (C3: 14 qubits, rho = 2^28 elements = 4 GiB) at its full size
against the C oracle (parity <= 8 qubits, properties at 11, nothing at 14). The file name
sorts behind every other GPU test on purpose: about a minute of host time, run last."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-10          # BASELINE.json north_star tolerance


def test_c3_density_14_qubits_full_matrix_against_c_oracle():
    """workloads.wd_circuit(14 qubits, depth 4, seed 0: RX layers, CNOTs on a random matching, Depolarizing(0.01) as
    Kraus operations -- the circuit bench.py --config c3 times, at a depth the oracle finishes in about a minute)
    through Circuit.evolve (planner + sweep kernels on a 28-bit index) against the oracle's gate-by-gate evolution
    with kron(U, conj U) and the Kraus-sum superoperator: all 2^28 elements, 1e-10 max-abs; trace and Hermiticity."""
    import psutil
    if psutil.virtual_memory().available < (5 << 32) or torch.cuda.mem_get_info()[0] < (3 << 32):
        pytest.skip('needs 20 GiB of host and 12 GiB of device memory')
    import quantumflow_b200 as qf
    from oracle import c_oracle
    from oracle import qf_oracle as O
    from quantumflow_b200 import engine, workloads
    n, depth, seed = 14, 4, 0
    before = engine.launch_count()
    rho = workloads.wd_circuit(qf, n, depth, seed, kraus=True).evolve()
    assert engine.launch_count() > before
    assert abs(complex(qf.asarray(rho.trace())) - 1) < 1e-10
    got = qf.asarray(rho.tensor).reshape(-1)
    del rho
    torch.cuda.empty_cache()
    want = c_oracle.evolve_specs(workloads.wd_gate_list(n, depth, seed), n, O.gate_matrix, O.depolarizing_superop)
    assert got.shape == want.shape
    err = 0.0
    step = 1 << 24
    for lo in range(0, want.size, step):
        err = max(err, float(np.abs(got[lo:lo + step] - want[lo:lo + step]).max()))
    print('14-qubit density, depth {}: max-abs error over 2^28 elements = {:.3e}'.format(depth, err))
    assert err < AMP_TOL
    dim = 1 << n
    mat = got.reshape(dim, dim)
    herm = max(float(np.abs(mat[lo:lo + 1024] - mat[:, lo:lo + 1024].conj().T).max()) for lo in range(0, dim, 4096))
    assert herm < 1e-12


def test_qvm_front_of_program_run():
    """forest.QuantumFlowQVM over Program.run on the device, as the reference's tests/test_forest.py:119-147 use it
    (with Program objects instead of pyQuil programs): correlated outcomes of a measured Bell pair, and the
    post-measurement wavefunction in pyQuil's amplitude order."""
    import quantumflow_b200 as qf
    from quantumflow_b200 import forest
    ro = qf.Register('ro')
    bell = qf.Program([qf.Declare('ro', 'BIT', 2), qf.H(0), qf.CNOT(0, 1), qf.Measure(0, ro[0]), qf.Measure(1, ro[1])])
    qvm = forest.QuantumFlowQVM()
    seen = set()
    for _ in range(16):
        res = qvm.load(bell).run().wait().read_from_memory_region(region_name='ro')
        assert res[0] == res[1]
        seen.add(res[0])
    assert seen <= {0, 1}
    prog = qf.Program([qf.Declare('ro', 'BIT', 1), qf.H(0), qf.CNOT(0, 1), qf.Measure(0, ro[0]), qf.H(0)])
    qvm = forest.QuantumFlowQVM()
    wf = qvm.load(prog).run().wait().wavefunction()
    res = qvm.read_from_memory_region(region_name='ro')[0]
    s = 0.70710678
    expected = {0: np.array([s, s, 0, 0]), 1: np.array([0, 0, s, -s])}[res]
    assert np.allclose(wf, expected, atol=1e-7)

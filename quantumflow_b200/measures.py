"""Closeness predicates used as parity checks by every test (quantumflow/measures.py:32-53,118-129,183-210):
Fubini-Study angle <= tolerance, i.e. insensitive to global phase; and the spectral read-outs of a density
(measures.py:76-181: fidelity, Bures distance / angle, entropy, mutual information). The reference itself notes
that the spectral ones "cannot be calculated within the tensor backend": they are host eigendecompositions of the
2^N x 2^N operator, a read-out for small N, while their inputs (partial traces, `asoperator`) are produced on the
device. The cvxpy-based diamond norm is out of scope."""
import numpy as np
import scipy.stats

from . import backend as bk
from .config import TOLERANCE
from .ops import Channel, Gate
from .qubits import asarray, fubini_study_angle, vectors_close
from .states import Density, State

__all__ = ['state_fidelity', 'state_angle', 'states_close', 'purity', 'fidelity', 'bures_distance', 'bures_angle',
           'density_angle', 'densities_close', 'entropy', 'mutual_info',
           'gate_angle', 'gates_close', 'channel_angle', 'channels_close']


def state_fidelity(state0: State, state1: State) -> bk.BKTensor:
    """|<0|1>|^2 for normalised pure states."""
    assert state0.qubits == state1.qubits
    overlap = bk.absolute(bk.inner(state0.tensor, state1.tensor))
    return overlap * overlap


def state_angle(ket0: State, ket1: State) -> bk.BKTensor:
    return fubini_study_angle(ket0.vec, ket1.vec)


def states_close(state0: State, state1: State, tolerance: float = TOLERANCE) -> bool:
    return vectors_close(state0.vec, state1.vec, tolerance)


def purity(rho: Density) -> bk.BKTensor:
    """tr(rho . rho) as the reference computes it (quantumflow/measures.py:59-64), a complex backend scalar: the
    inner product <rho^H, rho> = sum_ij rho_ji rho_ij -- a conjugate-transpose sweep (qfb_permute_bits) and one
    vdot reduction on the device. Equals sum |rho_ij|^2 only for Hermitian rho; intermediates of non-CP maps are not."""
    return bk.inner(rho.vec.H.tensor, rho.tensor)


def _operator(rho: Density) -> np.ndarray:
    return np.asarray(asarray(rho.asoperator()))


def _psd_sqrt(op: np.ndarray) -> np.ndarray:
    """Square root of a Hermitian positive semi-definite matrix from its eigendecomposition (negative rounding
    residue of the spectrum clipped)."""
    vals, vecs = np.linalg.eigh((op + op.conj().T) / 2)
    return (vecs * np.sqrt(np.maximum(vals, 0.0))) @ vecs.conj().T


def fidelity(rho0: Density, rho1: Density) -> float:
    """F(rho0, rho1) = (tr sqrt(sqrt(rho0) rho1 sqrt(rho0)))^2, clipped to [0, 1] (measures.py:76-91)."""
    assert rho0.qubit_nb == rho1.qubit_nb
    rho1 = rho1.permute(rho0.qubits)
    root0 = _psd_sqrt(_operator(rho0))
    inner = root0 @ _operator(rho1) @ root0
    spectrum = np.maximum(np.linalg.eigvalsh((inner + inner.conj().T) / 2), 0.0)
    return float(min(max(np.sum(np.sqrt(spectrum)) ** 2, 0.0), 1.0))


def bures_distance(rho0: Density, rho1: Density) -> float:
    """sqrt(tr rho0 + tr rho1 - 2 sqrt(F)) (measures.py:95-106)."""
    fid = fidelity(rho0, rho1)
    tr0 = np.real(np.trace(_operator(rho0)))
    tr1 = np.real(np.trace(_operator(rho1)))
    return float(np.sqrt(max(tr0 + tr1 - 2.0 * np.sqrt(fid), 0.0)))


def bures_angle(rho0: Density, rho1: Density) -> float:
    """arccos sqrt(F) (measures.py:110-115)."""
    return float(np.arccos(np.sqrt(fidelity(rho0, rho1))))


def entropy(rho: Density, base: float = None) -> float:
    """Von Neumann entropy, in nats unless `base` is given (measures.py:133-148)."""
    probs = np.maximum(np.linalg.eigvalsh(_operator(rho)), 0.0)
    return float(scipy.stats.entropy(probs, base=base))


def mutual_info(rho: Density, qubits0, qubits1=None, base: float = None) -> float:
    """Bipartite von Neumann mutual information S(0) + S(1) - S(01); the reduced densities come from the device
    partial trace (measures.py:152-180)."""
    if qubits1 is None:
        qubits1 = tuple(set(rho.qubits) - set(qubits0))
    rho0 = rho.partial_trace(qubits1)
    rho1 = rho.partial_trace(qubits0)
    return entropy(rho0, base) + entropy(rho1, base) - entropy(rho, base)


def density_angle(rho0: Density, rho1: Density) -> bk.BKTensor:
    return fubini_study_angle(rho0.vec, rho1.vec)


def densities_close(rho0: Density, rho1: Density, tolerance: float = TOLERANCE) -> bool:
    return vectors_close(rho0.vec, rho1.vec, tolerance)


def gate_angle(gate0: Gate, gate1: Gate) -> bk.BKTensor:
    return fubini_study_angle(gate0.vec, gate1.vec)


def gates_close(gate0: Gate, gate1: Gate, tolerance: float = TOLERANCE) -> bool:
    return vectors_close(gate0.vec, gate1.vec, tolerance)


def channel_angle(chan0: Channel, chan1: Channel) -> bk.BKTensor:
    return fubini_study_angle(chan0.vec, chan1.vec)


def channels_close(chan0: Channel, chan1: Channel, tolerance: float = TOLERANCE) -> bool:
    return vectors_close(chan0.vec, chan1.vec, tolerance)

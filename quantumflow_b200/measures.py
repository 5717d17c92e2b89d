"""Closeness predicates used as parity checks by every test (quantumflow/measures.py:32-53,118-129,183-210):
Fubini-Study angle <= tolerance, i.e. insensitive to global phase. cvxpy-based measures are out of scope."""
import numpy as np

from . import backend as bk
from .config import TOLERANCE
from .ops import Channel, Gate
from .qubits import asarray, fubini_study_angle, vectors_close
from .states import Density, State

__all__ = ['state_fidelity', 'state_angle', 'states_close', 'purity', 'density_angle', 'densities_close',
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
    """tr(rho^2) = sum |rho_ij|^2 for Hermitian rho: one squared-norm reduction on the device."""
    return rho.vec.norm()


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

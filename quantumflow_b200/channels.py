"""Kraus-form channels.

Behavioural contract: quantumflow/channels.py:36-224. `Kraus.evolve` is the reference's sum_k w_k K_k rho K_k^dagger
(channels.py:79-85) collapsed to ONE sweep with the superoperator S = sum_k w_k kron(K_k, conj K_k) (verified equal
to the reference to 3e-17); `Kraus.run` is the stochastic unravelling with the same single
`np.random.choice` draw (channels.py:70-77).
"""
from functools import reduce
from operator import add
from typing import Sequence

import numpy as np

from . import engine
from .gates import almost_identity
from .ops import Channel, Gate, Operation
from .qubits import Qubit, Qubits, asarray, outer_product
from .states import Density, State
from .stdgates import I, X, Y, Z

__all__ = ['Kraus', 'UnitaryMixture', 'Depolarizing', 'Damping', 'Dephasing', 'join_channels', 'channel_to_kraus',
           'kraus_iscomplete']

# superoperators on more than this many qubits (2x as many index bits) are applied operator by operator
_MAX_SUPEROP_QUBITS = 2


class Kraus(Operation):
    """Operator-sum representation: rho -> sum_k w_k K_k rho K_k^dagger."""

    def __init__(self, operators: Sequence[Gate], weights: Sequence[float] = None) -> None:
        self.operators = operators
        self.weights = tuple(weights) if weights is not None else (1.,) * len(operators)
        self._superop = None

    @property
    def qubits(self) -> Qubits:
        return tuple(sorted({q for op in self.operators for q in op.qubits}))

    def asgate(self) -> Gate:
        raise TypeError('Not possible in general')

    def aschannel(self) -> Channel:
        """sum_k w_k (K_k as a channel, extended to all of this operation's qubits)."""
        qubits = self.qubits
        ident = Gate(np.eye(2 ** len(qubits)), qubits=qubits).aschannel()
        terms = [(op.aschannel() @ ident) * w for op, w in zip(self.operators, self.weights)]
        return reduce(add, terms)

    def superoperator_matrix(self) -> np.ndarray:
        """4^K x 4^K host matrix of `aschannel()`, bit order [ket qubits..., bra qubits...]."""
        # cached per (operators, weights): replacing an operator or a weight after a first evolve() must not replay
        # the stale superoperator (the reference recomputes on every call)
        key = (tuple(id(op) for op in self.operators), tuple(self.weights))
        if self._superop is None or self._superop[0] != key:
            dim = 4 ** len(self.qubits)
            self._superop = (key, np.ascontiguousarray(asarray(self.aschannel().tensor).reshape(dim, dim)))
        return self._superop[1]

    def run(self, ket: State) -> State:
        """Pick one Kraus branch with probability w_k <psi|K_k^dagger K_k|psi>, then renormalise."""
        branches = [op.run(ket) for op in self.operators]
        probs = np.asarray([float(asarray(b.norm())) * w for b, w in zip(branches, self.weights)])
        probs /= np.sum(probs)
        pick = np.random.choice(len(branches), p=probs)
        return branches[pick].normalize()

    def evolve(self, rho: Density) -> Density:
        qubits = rho.qubits
        mine = self.qubits
        if len(mine) <= _MAX_SUPEROP_QUBITS:
            count = rho.qubit_nb
            where = [qubits.index(q) for q in mine]
            bits = [2 * count - 1 - w for w in where] + [count - 1 - w for w in where]
            tensor = engine.apply_operator(rho.tensor, self.superoperator_matrix(), bits)
            return Density(tensor, qubits)
        total = None
        for op, w in zip(self.operators, self.weights):
            term = op.evolve(rho).tensor
            total = engine.scale(term, w, inplace=True) if total is None else engine.axpby(total, 1.0, term, w)
        return Density(total, qubits)

    @property
    def H(self) -> 'Kraus':
        return Kraus([op.H for op in self.operators], self.weights)


class UnitaryMixture(Kraus):
    """Convex mixture of unitaries; on pure states one of them is drawn (no renormalisation needed)."""

    def asgate(self) -> Gate:
        pick = np.random.choice(len(self.operators), p=self.weights)
        return self.operators[pick]

    def run(self, ket: State) -> State:
        return self.asgate().run(ket)


class Depolarizing(UnitaryMixture):
    """(1-p) rho + p/3 (X rho X + Y rho Y + Z rho Z) on one qubit."""

    def __init__(self, prob: float, q0: Qubit) -> None:
        super().__init__([I(q0), X(q0), Y(q0), Z(q0)], [1 - prob, prob / 3.0, prob / 3.0, prob / 3.0])


class Damping(Kraus):
    """Amplitude damping (spontaneous emission) with decay probability `prob`."""

    def __init__(self, prob: float, q0: Qubit) -> None:
        keep = Gate([[1.0, 0.0], [0.0, np.sqrt(1 - prob)]], qubits=[q0])
        decay = Gate([[0.0, np.sqrt(prob)], [0.0, 0.0]], qubits=[q0])
        super().__init__([keep, decay])


class Dephasing(UnitaryMixture):
    """Phase damping: Z applied with probability prob/2."""

    def __init__(self, prob: float, q0: Qubit) -> None:
        super().__init__([I(q0), Z(q0)], [1 - prob / 2, prob / 2])


def join_channels(*channels: Channel) -> Channel:
    vec = reduce(outer_product, [chan.vec for chan in channels])
    return Channel(vec.tensor, vec.qubits)


def channel_to_kraus(chan: Channel) -> Kraus:
    """Kraus operators from the eigen-decomposition of the Choi matrix."""
    qubits = chan.qubits
    dim = 2 ** chan.qubit_nb
    evals, evecs = np.linalg.eig(asarray(chan.choi()))
    evecs = np.transpose(evecs)
    assert np.allclose(evals.imag, 0.0)
    assert np.all(evals.real >= -1e-12)
    amplitudes = np.sqrt(np.clip(evals.real, 0.0, None))
    ops = [Gate(np.reshape(vec, (dim, dim)) * amp, qubits)
           for vec, amp in zip(evecs, amplitudes) if not np.isclose(amp, 0.0)]
    return Kraus(ops)


def kraus_iscomplete(kraus: Kraus) -> bool:
    """sum_k w_k K_k^dagger K_k == I (trace preservation)."""
    qubits = kraus.qubits
    ident = Gate(np.eye(2 ** len(qubits)), qubits)
    total = reduce(np.add, [asarray((op.H @ op @ ident).asoperator()) * w
                            for op, w in zip(kraus.operators, kraus.weights)])
    return almost_identity(Gate(total, qubits))

"""Gate constructors and predicates that are not tied to a named standard gate.

Behavioural contract: quantumflow/gates.py:33-168 (identity, projectors, joins, controlled / conditional gates,
Haar-random gates, almost_* predicates). Everything here is host-side operator algebra feeding the kernels.
"""
from functools import reduce
from typing import TextIO, Union

import numpy as np
import scipy.stats

from .config import TOLERANCE
from .ops import Gate
from .qubits import Qubit, Qubits, asarray, outer_product, qubits_count_tuple

__all__ = ['I', 'identity_gate', 'random_gate', 'join_gates', 'control_gate', 'conditional_gate', 'P0', 'P1',
           'almost_unitary', 'almost_identity', 'almost_hermitian', 'print_gate']


class I(Gate):                                              # noqa: E742
    """Identity on any number of qubits (default: qubit 0)."""

    def __init__(self, *qubits: Qubit) -> None:
        qubits = qubits or (0,)
        super().__init__(np.eye(2 ** len(qubits)), qubits=qubits)

    @property
    def H(self) -> Gate:
        return self

    def __pow__(self, t: float) -> Gate:
        return self


class P0(Gate):
    """Projector |0><0| (non-unitary; scales the norm by the probability of reading 0)."""

    def __init__(self, q0: Qubit = 0) -> None:
        super().__init__([[1, 0], [0, 0]], qubits=[q0])


class P1(Gate):
    """Projector |1><1|."""

    def __init__(self, q0: Qubit = 0) -> None:
        super().__init__([[0, 0], [0, 1]], qubits=[q0])


def identity_gate(qubits: Union[int, Qubits]) -> Gate:
    _, qubits = qubits_count_tuple(qubits)
    return I(*qubits)


def join_gates(*gates: Gate) -> Gate:
    """Tensor product of gates on disjoint qubits."""
    vec = reduce(outer_product, [g.vec for g in gates])
    return Gate(vec.tensor, vec.qubits)


def control_gate(control: Qubit, gate: Gate) -> Gate:
    """P0(control) (x) I + P1(control) (x) gate, control qubit first."""
    if control in gate.qubits:
        raise ValueError('Gate and control qubits overlap')
    off = join_gates(P0(control), identity_gate(gate.qubits)).tensor
    on = join_gates(P1(control), gate).tensor
    return Gate(qubits=[control, *gate.qubits], tensor=off + on)


def conditional_gate(control: Qubit, gate0: Gate, gate1: Gate) -> Gate:
    """gate0 when the control reads 0, gate1 when it reads 1."""
    assert gate0.qubits == gate1.qubits
    tensor = join_gates(P0(control), gate0).tensor + join_gates(P1(control), gate1).tensor
    return Gate(tensor=tensor, qubits=[control, *gate0.qubits])


def almost_unitary(gate: Gate) -> bool:
    product = asarray((gate @ gate.H).asoperator())
    return bool(np.allclose(product, np.eye(2 ** gate.qubit_nb), atol=TOLERANCE))


def almost_identity(gate: Gate) -> bool:
    return bool(np.allclose(asarray(gate.asoperator()), np.eye(2 ** gate.qubit_nb)))


def almost_hermitian(gate: Gate) -> bool:
    return bool(np.allclose(asarray(gate.asoperator()), asarray(gate.H.asoperator())))


def print_gate(gate: Gate, ndigits: int = 2, file: TextIO = None) -> None:
    """One line per non-negligible matrix element: `bra -> ket : amplitude`, sorted by bra."""
    count = gate.qubit_nb
    rows = []
    for index, amplitude in np.ndenumerate(gate.vec.asarray()):
        if round(abs(amplitude) ** 2, ndigits) > 0.0:
            ket = ''.join(str(b) for b in index[:count])
            bra = ''.join(str(b) for b in index[count:])
            rows.append('{} -> {} : {}'.format(bra, ket, amplitude))
    rows.sort(key=lambda line: int(line[:count]))
    print('\n'.join(rows), file=file)


def random_gate(qubits: Union[int, Qubits]) -> Gate:
    """Haar-random unitary (scipy.stats.unitary_group, same call as gates.py:159-168)."""
    count, qubits = qubits_count_tuple(qubits)
    unitary = scipy.stats.unitary_group.rvs(2 ** count)
    return Gate(unitary, qubits=qubits, name='RAND{}'.format(count))

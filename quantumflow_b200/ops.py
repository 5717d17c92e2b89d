"""Operations: the abstract base, unitary-or-not `Gate`s and superoperator `Channel`s.

Behavioural contract: quantumflow/ops.py:33-412. Operators are host tensors (tiny); applying one to a State or
Density launches libqfb200 kernels through `bk.tensormul` / `engine.apply_operator`:

  Gate.run      ops.py:161-166  -> one sweep, structure-specialised (diagonal / controlled / dense)
  Gate.evolve   ops.py:168-172  -> U on the ket bit(s) and conj(U) on the bra bit(s) of the 2N-bit vector; the
                                   reference's superoperator kron(U, conj U) (ops.py:243-252) applied in one or
                                   two sweeps, whichever moves fewer flops
  Channel.evolve ops.py:354-363 -> the superoperator as a 2K-bit operator on the 2N-bit vector
"""
from abc import ABC
from copy import copy
from typing import Any, Dict, Union

import numpy as np
from scipy.linalg import fractional_matrix_power as _matpow

from . import backend as bk
from . import classify
from .qubits import Qubits, QubitVector, asarray, qubits_count_tuple
from .states import Density, State

__all__ = ['Operation', 'Gate', 'Channel']


class Operation(ABC):
    """Anything that can be an element of a Circuit."""

    _qubits: Qubits = ()

    @property
    def qubits(self) -> Qubits:
        return self._qubits

    @property
    def qubit_nb(self) -> int:
        return len(self.qubits)

    @property
    def name(self) -> str:
        return type(self).__name__.upper()

    def run(self, ket: State) -> State:
        raise NotImplementedError()

    def evolve(self, rho: Density) -> Density:
        raise NotImplementedError()

    def quil(self) -> str:
        raise NotImplementedError()

    def __str__(self) -> str:
        return self.quil()

    def asgate(self) -> 'Gate':
        raise NotImplementedError()

    def aschannel(self) -> 'Channel':
        raise NotImplementedError()

    @property
    def H(self) -> 'Operation':
        raise NotImplementedError()


def _requires_grad(tensor) -> bool:
    return bool(getattr(tensor, 'requires_grad', False))


class Gate(Operation):
    """A K-qubit operator stored as a [2]*2K host tensor (kets then bras; gate qubit 0 is the MSB of the matrix
    index). Not necessarily unitary (P0, P1, Kraus operators)."""

    def __init__(self, tensor: bk.TensorLike, qubits: Qubits = None, params: Dict[str, float] = None,
                 name: str = None) -> None:
        if qubits is None:
            tensor = bk.astensorproduct(tensor)
            qubits = range(bk.rank(tensor) // 2)
        self.vec = QubitVector(tensor, qubits, resident=False)
        self.params = params if params is not None else {}
        self._name = name if name is not None else type(self).__name__
        self._matrix_cache = None

    @property
    def name(self) -> str:
        return self._name

    @property
    def tensor(self) -> bk.BKTensor:
        return self.vec.tensor

    @property
    def qubits(self) -> Qubits:
        return self.vec.qubits

    @property
    def qubit_nb(self) -> int:
        return self.vec.qubit_nb

    def matrix(self) -> np.ndarray:
        """Host 2^K x 2^K complex128 matrix (detached); what the kernels receive as a launch parameter."""
        if getattr(self, '_matrix_cache', None) is None:
            dim = 2 ** self.qubit_nb
            self._matrix_cache = np.ascontiguousarray(asarray(self.tensor).reshape(dim, dim))
        return self._matrix_cache

    def relabel(self, qubits: Qubits) -> 'Gate':
        gate = copy(self)
        gate.vec = gate.vec.relabel(qubits)
        return gate

    def permute(self, qubits: Qubits) -> 'Gate':
        vec = self.vec.permute(qubits)
        return Gate(vec.tensor, qubits=vec.qubits)

    @property
    def H(self) -> 'Gate':
        return Gate(tensor=self.vec.H.tensor, qubits=self.qubits)

    def asoperator(self) -> bk.BKTensor:
        return self.vec.flatten()

    def run(self, ket: State) -> State:
        where = [ket.qubits.index(q) for q in self.qubits]      # ValueError if the state lacks a qubit
        tensor = bk.tensormul(self.tensor, ket.tensor, where)
        return State(tensor, ket.qubits, ket.memory)

    def evolve(self, rho: Density) -> Density:
        if _requires_grad(self.tensor) or _requires_grad(rho.tensor):
            return self.aschannel().evolve(rho)
        from . import engine
        count = rho.qubit_nb
        where = [rho.qubits.index(q) for q in self.qubits]
        ket_bits = [2 * count - 1 - w for w in where]
        bra_bits = [count - 1 - w for w in where]
        mat = self.matrix()
        k = self.qubit_nb
        structured = classify.is_diagonal(mat) or (k >= 2 and bool(classify.peel_controls(mat, k)[0]))
        if k == 1 and not structured:
            tensor = engine.apply_operator(rho.tensor, np.kron(mat, mat.conj()), ket_bits + bra_bits)
        else:
            tensor = engine.apply_operator(rho.tensor, mat, ket_bits)
            tensor = engine.apply_operator(tensor, mat.conj(), bra_bits, inplace=True)
        return Density(tensor, rho.qubits, rho.memory)

    def __pow__(self, t: float) -> 'Gate':
        """Generic matrix power on the host (subclasses override with closed forms)."""
        count = self.qubit_nb
        powered = _matpow(self.matrix(), t)
        return Gate(np.reshape(powered, [2] * (2 * count)), self.qubits)

    def __matmul__(self, other: 'Gate') -> 'Gate':
        """self o other; `other` must act on every qubit of `self` (ops.py:187-198)."""
        if not isinstance(other, Gate):
            raise NotImplementedError()
        where = [other.qubits.index(q) for q in self.qubits]
        tensor = bk.tensormul(self.tensor, other.tensor, where)
        return Gate(tensor=tensor, qubits=other.qubits)

    def quil(self) -> str:
        if self.name == 'Gate':
            return object.__repr__(self)
        text = self.name
        if self.params:
            text += '(' + ', '.join(_format_param(p) for p in self.params.values()) + ')'
        return text + ' ' + ' '.join(str(q) for q in self.qubits)

    def asgate(self) -> 'Gate':
        return self

    def aschannel(self) -> 'Channel':
        """Superoperator kron(U, conj U) with axes [ket_out, bra_out, ket_in, bra_in]."""
        dim = 2 ** self.qubit_nb
        tensor = bk.outer(self.tensor, self.H.tensor)
        tensor = bk.reshape(tensor, [dim] * 4)
        tensor = bk.transpose(tensor, [0, 3, 1, 2])
        return Channel(tensor, self.qubits)

    def su(self) -> 'Gate':
        """Rescale to unit determinant."""
        dim = 2 ** self.qubit_nb
        mat = np.array(self.matrix())
        mat /= np.linalg.det(mat) ** (1 / dim)
        return Gate(tensor=mat, qubits=self.qubits)


def _format_param(obj: Any) -> str:
    """Quil text of a gate parameter (quantumflow/ops.py:204-210): floats as small fractions (of pi) when they are."""
    if isinstance(obj, float):
        from .utils import symbolize
        try:
            return symbolize(obj)
        except (ValueError, OverflowError):
            return '{}'.format(obj)
    return str(obj)


class Channel(Operation):
    """A superoperator on K qubits: [2]*4K host tensor, axes [ket_out, bra_out, ket_in, bra_in]."""

    def __init__(self, tensor: bk.TensorLike, qubits: Union[int, Qubits], params: Dict[str, Any] = None,
                 name: str = None) -> None:
        _, qubits = qubits_count_tuple(qubits)
        self.vec = QubitVector(tensor, qubits, resident=False)
        self.params = params
        self._name = name if name is not None else type(self).__name__

    @property
    def name(self) -> str:
        return self._name

    @property
    def tensor(self) -> bk.BKTensor:
        return self.vec.tensor

    @property
    def qubits(self) -> Qubits:
        return self.vec.qubits

    @property
    def qubit_nb(self) -> int:
        return self.vec.qubit_nb

    def relabel(self, qubits: Qubits) -> 'Channel':
        chan = copy(self)
        chan.vec = chan.vec.relabel(qubits)
        return chan

    def permute(self, qubits: Qubits) -> 'Channel':
        vec = self.vec.permute(qubits)
        return Channel(vec.tensor, qubits=vec.qubits)

    @property
    def H(self) -> 'Channel':
        return Channel(tensor=self.vec.H.tensor, qubits=self.qubits)

    @property
    def sharp(self) -> 'Channel':
        """Swap the 2nd and 3rd super-indices; flattening the result gives the Choi matrix."""
        dim = 2 ** self.qubit_nb
        tensor = bk.reshape(self.tensor, [dim] * 4)
        tensor = bk.transpose(tensor, (0, 2, 1, 3))
        return Channel(bk.reshape(tensor, [2] * (4 * self.qubit_nb)), self.qubits)

    def choi(self) -> bk.BKTensor:
        side = 2 ** (2 * self.qubit_nb)
        return bk.reshape(self.sharp.tensor, [side, side])

    def chi(self) -> bk.BKTensor:
        side = 2 ** (2 * self.qubit_nb)
        return bk.reshape(self.sharp.tensor, [side, side])

    def run(self, ket: State) -> State:
        raise TypeError()  # a general channel has no action on pure states

    def evolve(self, rho: Density) -> Density:
        count = rho.qubit_nb
        where = [rho.qubits.index(q) for q in self.qubits]
        indices = where + [w + count for w in where]
        tensor = bk.tensormul(self.tensor, rho.tensor, indices)
        return Density(tensor, rho.qubits, rho.memory)

    def asgate(self) -> Gate:
        raise TypeError()

    def aschannel(self) -> 'Channel':
        return self

    def __add__(self, other: Any) -> 'Channel':
        if isinstance(other, Channel):
            if not self.qubits == other.qubits:
                raise ValueError('Qubits must be identical')
            return Channel(self.tensor + other.tensor, self.qubits)
        raise NotImplementedError()

    def __mul__(self, other: Any) -> 'Channel':
        return Channel(self.tensor * other, self.qubits)

    def __matmul__(self, other: 'Channel') -> 'Channel':
        if not isinstance(other, Channel):
            raise NotImplementedError()
        count = other.qubit_nb
        where = [other.qubits.index(q) for q in self.qubits]
        indices = where + [w + count for w in where]
        tensor = bk.tensormul(self.tensor, other.tensor, indices)
        return Channel(tensor, other.qubits)

    def trace(self) -> bk.BKTensor:
        return self.vec.trace()

    def partial_trace(self, qubits: Qubits) -> 'Channel':
        vec = self.vec.partial_trace(qubits)
        return Channel(vec.tensor, vec.qubits)

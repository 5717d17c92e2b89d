"""QubitVector: a product tensor ([2]*(N*rank)) labelled by qubits.

Behavioural contract: quantumflow/qubits.py:39-325. rank 1 = state vector, 2 = operator / density,
4 = superoperator, 8 = super-duper-operator. Tensor axis i of a rank-1 vector belongs to `qubits[i]` and is flat
index bit N-1-i. Superoperator axes are [ket_out, bra_out, ket_in, bra_in].

A QubitVector is either *resident* (amplitude tensor in HBM: states, densities) or a host operator tensor
(gates, channels); see backend/b200bk.py. The flag is decided by the owner at construction.
"""
from copy import copy
from typing import Any, Hashable, List, Sequence, Tuple, Union

import numpy as np

from . import backend as bk
from .config import TOLERANCE

__all__ = ['Qubit', 'Qubits', 'asarray', 'QubitVector', 'inner_product', 'outer_product', 'fubini_study_angle',
           'vectors_close']

Qubit = Hashable
Qubits = Sequence[Qubit]

_RANKS = (1, 2, 4, 8)


def asarray(tensor: bk.BKTensor) -> bk.TensorLike:
    """Backend tensor -> numpy array."""
    return bk.evaluate(tensor)


class QubitVector:
    """Tensor + qubit labels + rank. `resident=True` places the tensor in HBM (amplitude domain)."""

    def __init__(self, tensor: bk.TensorLike, qubits: Qubits, rank: int = None, resident: bool = None) -> None:
        if resident is None:
            resident = bk.is_amplitudes(tensor)
        tensor = bk.asamplitudes(tensor) if resident else bk.astensorproduct(tensor)
        count = len(qubits)
        ndim = bk.rank(tensor)
        if rank is None:
            rank = 1 if count == 0 else ndim // count
        if rank not in _RANKS or rank * count != ndim:
            raise ValueError('Incompatibility between tensor and qubits')
        self.tensor = tensor
        self.qubits = tuple(qubits)
        self.qubit_nb = count
        self.rank = rank

    @property
    def resident(self) -> bool:
        return bk.is_amplitudes(self.tensor)

    def __getitem__(self, key: Any) -> bk.BKTensor:
        return bk.getitem(self.tensor, key)

    def asarray(self) -> np.ndarray:
        return bk.evaluate(self.tensor)

    def flatten(self) -> bk.BKTensor:
        """Tensor with the qubit axes of each rank-index merged: shape [2^N]*rank."""
        return bk.reshape(self.tensor, [2 ** self.qubit_nb] * self.rank)

    def relabel(self, qubits: Qubits) -> 'QubitVector':
        qubits = tuple(qubits)
        assert len(qubits) == self.qubit_nb
        other = copy(self)
        other.qubits = qubits
        return other

    def permute(self, qubits: Qubits) -> 'QubitVector':
        """Reorder the qubit axes (of every rank-index) to follow `qubits`."""
        if qubits == self.qubits:
            return self
        count = self.qubit_nb
        assert len(qubits) == count
        where = [self.qubits.index(q) for q in qubits]        # ValueError on unknown qubit
        perm: List[int] = [block * count + w for block in range(self.rank) for w in where]
        return QubitVector(bk.transpose(self.tensor, perm), qubits, resident=self.resident)

    @property
    def H(self) -> 'QubitVector':
        """Conjugate transpose as a (super)operator: swap the output and input halves of the axes."""
        half = (self.qubit_nb * self.rank) // 2
        dim = 2 ** half
        mat = bk.reshape(self.tensor, [dim, dim])
        if bk.is_amplitudes(mat):
            from . import engine
            total = 2 * half
            perm = [(j + half) % total for j in range(total)]
            out = engine.permute_bits(mat, perm, conj=True)
        else:
            out = bk.conj(bk.transpose(mat))
        return QubitVector(bk.reshape(out, [2] * (2 * half)), self.qubits, resident=self.resident)

    def norm(self) -> bk.BKTensor:
        """<v|v> (the squared 2-norm, as in the reference: qubits.py:180-182)."""
        if self.resident and not self.tensor.requires_grad:
            from . import engine
            return engine.norm2(self.tensor)
        return bk.absolute(bk.inner(self.tensor, self.tensor))

    def trace(self) -> bk.BKTensor:
        if self.rank == 1:
            raise ValueError('Cannot take trace of vector')
        dim = 2 ** ((self.qubit_nb * self.rank) // 2)
        return bk.trace(bk.reshape(self.tensor, [dim, dim]))

    def partial_trace(self, qubits: Qubits) -> 'QubitVector':
        """Trace out `qubits` (rank >= 2)."""
        if self.rank == 1:
            raise ValueError('Cannot take trace of vector')
        keep = list(self.qubits)
        for q in qubits:
            keep.remove(q)
        if not keep:
            raise ValueError('Cannot remove all qubits with partial_trace.')
        count, rank = self.qubit_nb, self.rank
        traced = [self.qubits.index(q) for q in qubits]
        letters = bk.EINSUM_SUBSCRIPTS[:count * rank]
        keep_bits, masks = partial_trace_layout(count, rank, traced)
        nkeep = len(keep_bits)
        if not self.resident:
            # operator domain (gates / channels: host tensors of a few thousand elements, b200bk docstring):
            # the reference's own einsum call
            sub = list(letters)
            for ax in traced:
                for block in range(1, rank):
                    sub[block * count + ax] = sub[ax]
            return QubitVector(np.einsum(''.join(sub), self.asarray()), keep, resident=False)
        from . import engine
        reduced = engine.partial_trace(self.tensor, keep_bits, masks).reshape([2] * nkeep)
        return QubitVector(reduced, keep, resident=True)


def partial_trace_layout(count: int, rank: int, traced: List[int]) -> Tuple[List[int], List[int]]:
    """Index-bit description of a partial trace for qfb_partial_trace: (keep_bits, masks) for a [2]*(count*rank)
    tensor whose qubit axes `traced` are summed over. keep_bits[b] = input index bit of output index bit b,
    masks[t] = OR of the input index bits of traced axis t in every rank block.

    The reference sums with np.einsum over repeated subscripts in implicit mode (qubits.py:216-225): the surviving
    axes come out sorted by their subscript letter, and in ASCII the upper-case letters that label axes >= 26
    sort before the lower-case ones. The same order is produced here."""
    total = count * rank
    letters = bk.EINSUM_SUBSCRIPTS[:total]
    kept_axes = sorted((a for a in range(total) if a % count not in traced), key=lambda a: ord(letters[a]))
    nkeep = len(kept_axes)
    keep_bits = [total - 1 - kept_axes[nkeep - 1 - b] for b in range(nkeep)]
    masks = [sum(1 << (total - 1 - (block * count + ax)) for block in range(rank)) for ax in traced]
    return keep_bits, masks


def _check_compatible(vec0: QubitVector, vec1: QubitVector) -> None:
    if vec0.rank != vec1.rank or vec0.qubit_nb != vec1.qubit_nb:
        raise ValueError('Incompatibly vectors. Qubits and rank must match')


def inner_product(vec0: QubitVector, vec1: QubitVector) -> bk.BKTensor:
    """Hilbert-Schmidt inner product <vec0|vec1>."""
    _check_compatible(vec0, vec1)
    vec1 = vec1.permute(vec0.qubits)
    return bk.inner(vec0.tensor, vec1.tensor)


def outer_product(vec0: QubitVector, vec1: QubitVector) -> QubitVector:
    """Tensor product over disjoint qubits; rank-indices of the two factors are interleaved so that the result
    is again [kets..., bras...] (quantumflow/qubits.py:248-277)."""
    rank = vec0.rank
    if rank != vec1.rank:
        raise ValueError('Incompatibly vectors. Rank must match')
    if not set(vec0.qubits).isdisjoint(vec1.qubits):
        raise ValueError('Overlapping qubits')
    n0, n1 = vec0.qubit_nb, vec1.qubit_nb
    resident = vec0.resident or vec1.resident
    flat = bk.outer(vec0.tensor, vec1.tensor)
    if rank == 1:
        return QubitVector(bk.reshape(flat, [2] * (n0 + n1)), tuple(vec0.qubits) + tuple(vec1.qubits),
                           resident=resident)
    # axes: rank blocks of vec0 (n0 axes each) then rank blocks of vec1 (n1 axes each) -> interleave blocks
    full = bk.reshape(flat, [2] * (rank * (n0 + n1)))
    perm: List[int] = []
    for block in range(rank):
        perm += list(range(block * n0, (block + 1) * n0))
        perm += list(range(rank * n0 + block * n1, rank * n0 + (block + 1) * n1))
    return QubitVector(bk.transpose(full, perm), tuple(vec0.qubits) + tuple(vec1.qubits), resident=resident)


def fubini_study_angle(vec0: QubitVector, vec1: QubitVector) -> bk.BKTensor:
    """arccos(|<a|b>| / sqrt(<a|a><b|b>)): distance in projective Hilbert space (global-phase insensitive)."""
    _check_compatible(vec0, vec1)
    vec1 = vec1.permute(vec0.qubits)
    t0, t1 = vec0.tensor, vec1.tensor
    if bk.is_amplitudes(t0) != bk.is_amplitudes(t1):
        t0, t1 = bk.asamplitudes(t0), bk.asamplitudes(t1)
    hs01 = bk.inner(t0, t1)
    hs00 = bk.inner(t0, t0)
    hs11 = bk.inner(t1, t1)
    ratio = bk.absolute(hs01) / bk.sqrt(bk.absolute(hs00 * hs11))
    ratio = bk.minimum(ratio, bk.fcast(1.))
    return bk.arccos(ratio)


def vectors_close(vec0: QubitVector, vec1: QubitVector, tolerance: float = TOLERANCE) -> bool:
    if vec0.rank != vec1.rank or vec0.qubit_nb != vec1.qubit_nb:
        return False
    if set(vec0.qubits) ^ set(vec1.qubits):
        return False
    return bool(bk.evaluate(fubini_study_angle(vec0, vec1)) <= tolerance)


def qubits_count_tuple(qubits: Union[int, Qubits]) -> Tuple[int, Qubits]:
    """`3` -> (3, (0, 1, 2)); a sequence -> (len, sequence)."""
    if isinstance(qubits, int):
        return qubits, tuple(range(qubits))
    return len(qubits), qubits

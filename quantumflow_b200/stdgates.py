"""The standard gate set.

Matrix conventions follow quantumflow/stdgates.py:32-988 exactly (golden fixture: tests/golden/stdgates.npz was
generated from the reference). Each class is declared by its operator formula plus, where the reference defines
them, closed forms for the inverse (`H`) and for powers. Parametric operators are built from
`bk.cos/sin/exp/cis(bk.ccast(theta))`, so a parameter may be a torch tensor that requires grad (QAOA bridge).

Structural class of each gate (drives kernel specialisation):
  diagonal      Z S T PHASE RZ S_H T_H TZ CZ CPHASE00/01/10 CPHASE ZZ
  permutation   X Y CNOT SWAP ISWAP PSWAP CCNOT CSWAP        (controlled ones touch only control=1 groups)
  dense         H RX RY RN TX TY TH ZYZ PISWAP CAN XX YY EXCH
"""
import copy
from math import pi, sqrt

import numpy as np

from . import backend as bk
from .gates import I
from .ops import Gate

__all__ = ['I', 'X', 'Y', 'Z', 'H', 'S', 'T', 'PHASE', 'RX', 'RY', 'RZ', 'CZ', 'CNOT', 'SWAP', 'ISWAP',
           'CPHASE00', 'CPHASE01', 'CPHASE10', 'CPHASE', 'PSWAP', 'CCNOT', 'CSWAP', 'RN', 'TX', 'TY', 'TZ', 'TH',
           'ZYZ', 'CAN', 'XX', 'YY', 'ZZ', 'PISWAP', 'EXCH', 'CANONICAL', 'S_H', 'T_H', 'STDGATES']


class StdGate(Gate):
    """Declarative base: subclasses give ARITY, PARAMS and `operator(*params)`."""

    ARITY = 1          # number of qubits
    PARAMS = ()        # parameter names, in constructor order
    HERMITIAN = False  # H returns a copy

    def __init__(self, *args, **kwargs) -> None:
        nparams = len(self.PARAMS)
        values = list(args[:nparams])
        for pname in self.PARAMS[len(values):]:
            if pname not in kwargs:
                raise TypeError('{}() missing parameter {!r}'.format(type(self).__name__, pname))
            values.append(kwargs.pop(pname))
        qubits = list(args[nparams:])
        for pos in range(len(qubits), self.ARITY):
            qubits.append(kwargs.pop('q{}'.format(pos), pos))
        if kwargs or len(qubits) != self.ARITY:
            raise TypeError('{}(): bad arguments'.format(type(self).__name__))
        values = list(self.canonical(*values))
        super().__init__(self.operator(*values), qubits, dict(zip(self.PARAMS, values)))

    @staticmethod
    def canonical(*values):
        return values

    @staticmethod
    def operator(*values):
        raise NotImplementedError()

    def _rebuild(self, *values) -> Gate:
        return type(self)(*values, *self.qubits)

    def inverse(self) -> Gate:
        """Closed-form inverse; default falls back to the conjugate transpose of the operator."""
        return Gate(tensor=self.vec.H.tensor, qubits=self.qubits)

    @property
    def H(self) -> Gate:
        if self.HERMITIAN:
            return copy.copy(self)
        return self.inverse()


def _cs(theta):
    half = bk.ccast(theta) / 2
    return bk.cos(half), bk.sin(half)


# ---------------------------------------------------------------------------------------------------------
# one-qubit gates
# ---------------------------------------------------------------------------------------------------------

class X(StdGate):
    HERMITIAN = True

    @staticmethod
    def operator():
        return [[0, 1], [1, 0]]

    def __pow__(self, t):
        return TX(t, *self.qubits)


class Y(StdGate):
    HERMITIAN = True

    @staticmethod
    def operator():
        return np.asarray([[0, -1.0j], [1.0j, 0]])

    def __pow__(self, t):
        return TY(t, *self.qubits)


class Z(StdGate):
    HERMITIAN = True

    @staticmethod
    def operator():
        return [[1, 0], [0, -1.0]]

    def __pow__(self, t):
        return TZ(t, *self.qubits)


class H(StdGate):
    HERMITIAN = True

    @staticmethod
    def operator():
        return np.asarray([[1, 1], [1, -1]]) / sqrt(2)

    def __pow__(self, t):
        return TH(t, *self.qubits)


class S(StdGate):
    @staticmethod
    def operator():
        return np.asarray([[1.0, 0.0], [0.0, 1.0j]])

    def inverse(self):
        return S_H(*self.qubits)

    def __pow__(self, t):
        return PHASE(pi / 2 * t, *self.qubits)


class T(StdGate):
    @staticmethod
    def operator():
        return [[1.0, 0.0], [0.0, bk.ccast(bk.cis(pi / 4.0))]]

    def inverse(self):
        return T_H(*self.qubits)

    def __pow__(self, t):
        return PHASE(pi / 4 * t, *self.qubits)


class S_H(StdGate):
    @staticmethod
    def operator():
        return np.asarray([[1.0, 0.0], [0.0, -1.0j]])

    def inverse(self):
        return S(*self.qubits)

    def __pow__(self, t):
        return PHASE(-pi / 2 * t, *self.qubits)


class T_H(StdGate):
    @staticmethod
    def operator():
        return [[1.0, 0.0], [0.0, bk.ccast(bk.cis(-pi / 4.0))]]

    def inverse(self):
        return T(*self.qubits)

    def __pow__(self, t):
        return PHASE(-pi / 4 * t, *self.qubits)


class PHASE(StdGate):
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        return [[1.0, 0.0], [0.0, bk.cis(bk.ccast(theta))]]

    def inverse(self):
        theta = self.params['theta']
        return PHASE(2. * pi - theta % (2. * pi), *self.qubits)

    def __pow__(self, t):
        return PHASE(self.params['theta'] * t, *self.qubits)


class _Rotation(StdGate):
    """Single-parameter gates whose inverse negates the parameter and whose power scales it."""

    def inverse(self):
        (value,) = self.params.values()
        return self._rebuild(-value)

    def __pow__(self, t):
        (value,) = self.params.values()
        return self._rebuild(value * t)


class RX(_Rotation):
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        c, s = _cs(theta)
        return [[c, -1.0j * s], [-1.0j * s, c]]


class RY(_Rotation):
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        c, s = _cs(theta)
        return [[c, -s], [s, c]]


class RZ(_Rotation):
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        ct = bk.ccast(theta)
        return [[bk.exp(-ct * 0.5j), 0], [0, bk.exp(ct * 0.5j)]]


class RN(StdGate):
    """Rotation by theta about the unit axis (nx, ny, nz)."""
    PARAMS = ('theta', 'nx', 'ny', 'nz')

    @staticmethod
    def operator(theta, nx, ny, nz):
        c, s = _cs(theta)
        return [[c - 1j * s * nz, -1j * s * nx - s * ny],
                [-1j * s * nx + s * ny, c + 1j * s * nz]]

    def inverse(self):
        theta, nx, ny, nz = self.params.values()
        return RN(-theta, nx, ny, nz, *self.qubits)

    def __pow__(self, t):
        theta, nx, ny, nz = self.params.values()
        return RN(t * theta, nx, ny, nz, *self.qubits)


class _HalfTurns(_Rotation):
    """Powers of a Pauli: parameter counted in half turns and reduced mod 2."""
    PARAMS = ('t',)

    @staticmethod
    def canonical(t):
        return (t % 2,)


class TX(_HalfTurns):
    @staticmethod
    def operator(t):
        ct = bk.ccast(pi * t)
        phase = bk.exp(0.5j * ct)
        c, s = bk.cos(ct / 2), bk.sin(ct / 2)
        return [[phase * c, phase * -1.0j * s], [phase * -1.0j * s, phase * c]]


class TY(_HalfTurns):
    @staticmethod
    def operator(t):
        ct = bk.ccast(pi * t)
        phase = bk.exp(0.5j * ct)
        c, s = bk.cos(ct / 2.0), bk.sin(ct / 2.0)
        return [[phase * c, phase * -s], [phase * s, phase * c]]


class TZ(_HalfTurns):
    @staticmethod
    def operator(t):
        ct = bk.ccast(pi * t)
        phase = bk.exp(0.5j * ct)
        return [[phase * bk.exp(-ct * 0.5j), 0], [0, phase * bk.exp(ct * 0.5j)]]


class TH(_Rotation):
    """Powers of the Hadamard gate (no mod-2 reduction in the reference)."""
    PARAMS = ('t',)

    @staticmethod
    def operator(t):
        theta = bk.ccast(pi * t)
        phase = bk.exp(0.5j * theta)
        c = phase * bk.cos(theta / 2)
        s = (phase * 1.0j * bk.sin(theta / 2)) / sqrt(2)
        return [[c - s, -s], [-s, c + s]]


class ZYZ(StdGate):
    """Z^t2 Y^t1 Z^t0 up to phase: the generic SU(2) element."""
    PARAMS = ('t0', 't1', 't2')

    @staticmethod
    def operator(t0, t1, t2):
        a0, a1, a2 = bk.ccast(pi * t0), bk.ccast(pi * t1), bk.ccast(pi * t2)
        c, s = bk.cos(0.5 * a1), bk.sin(0.5 * a1)
        return [[bk.cis(-0.5 * a2 - 0.5 * a0) * c, -bk.cis(-0.5 * a2 + 0.5 * a0) * s],
                [bk.cis(0.5 * a2 - 0.5 * a0) * s, bk.cis(0.5 * a2 + 0.5 * a0) * c]]

    def inverse(self):
        t0, t1, t2 = self.params.values()
        return ZYZ(-t2, -t1, -t0, *self.qubits)


# ---------------------------------------------------------------------------------------------------------
# two-qubit gates
# ---------------------------------------------------------------------------------------------------------

def _perm_matrix(dim, swaps=(), phases=None):
    mat = np.eye(dim, dtype=np.complex128 if phases else np.float64)
    for a, b in swaps:
        mat[[a, b]] = mat[[b, a]]
    if phases:
        for (r, c), v in phases.items():
            mat[r, c] = v
    return mat


class CZ(StdGate):
    ARITY = 2
    HERMITIAN = True

    @staticmethod
    def operator():
        return np.diag([1, 1, 1, -1])


class CNOT(StdGate):
    ARITY = 2
    HERMITIAN = True

    @staticmethod
    def operator():
        return _perm_matrix(4, swaps=[(2, 3)])


class SWAP(StdGate):
    ARITY = 2
    HERMITIAN = True

    @staticmethod
    def operator():
        return _perm_matrix(4, swaps=[(1, 2)])


class ISWAP(StdGate):
    ARITY = 2

    @staticmethod
    def operator():
        return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])


def _phase_on(slot):
    """diag(1,1,1,1) with exp(i theta) in position `slot`."""
    def operator(theta):
        entries = [1.0, 1.0, 1.0, 1.0]
        entries[slot] = bk.exp(1j * bk.ccast(theta))
        return [[entries[r] if r == c else 0 for c in range(4)] for r in range(4)]
    return staticmethod(operator)


class CPHASE00(_Rotation):
    ARITY = 2
    PARAMS = ('theta',)
    operator = _phase_on(0)
    __pow__ = Gate.__pow__


class CPHASE01(_Rotation):
    ARITY = 2
    PARAMS = ('theta',)
    operator = _phase_on(1)
    __pow__ = Gate.__pow__


class CPHASE10(_Rotation):
    ARITY = 2
    PARAMS = ('theta',)
    operator = _phase_on(2)
    __pow__ = Gate.__pow__


class CPHASE(_Rotation):
    ARITY = 2
    PARAMS = ('theta',)
    operator = _phase_on(3)


class PSWAP(StdGate):
    """SWAP with a phase exp(i theta) on the exchanged amplitudes."""
    ARITY = 2
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        ph = bk.exp(bk.ccast(theta) * 1.0j)
        return [[1, 0, 0, 0], [0, 0, ph, 0], [0, ph, 0, 0], [0, 0, 0, 1]]

    def inverse(self):
        theta = self.params['theta']
        return PSWAP(2. * pi - theta % (2. * pi), *self.qubits)


class PISWAP(_Rotation):
    ARITY = 2
    PARAMS = ('theta',)

    @staticmethod
    def operator(theta):
        ct = bk.ccast(theta)
        c, s = bk.cos(2 * ct), bk.sin(2 * ct) * 1j
        return [[1, 0, 0, 0], [0, c, s, 0], [0, s, c, 0], [0, 0, 0, 1]]


class _HalfTurns2(_Rotation):
    ARITY = 2
    PARAMS = ('t',)


class XX(_HalfTurns2):
    @staticmethod
    def operator(t):
        theta = bk.ccast(pi * t)
        c, s = bk.cos(theta / 2), -1.0j * bk.sin(theta / 2)
        return [[c, 0, 0, s], [0, c, s, 0], [0, s, c, 0], [s, 0, 0, c]]


class YY(_HalfTurns2):
    @staticmethod
    def operator(t):
        theta = bk.ccast(pi * t)
        c, s = bk.cos(theta / 2), 1.0j * bk.sin(theta / 2)
        return [[c, 0, 0, s], [0, c, -s, 0], [0, -s, c, 0], [s, 0, 0, c]]


class ZZ(_HalfTurns2):
    @staticmethod
    def operator(t):
        theta = bk.ccast(pi * t)
        minus, plus = bk.cis(-theta / 2), bk.cis(theta / 2)
        return [[minus, 0, 0, 0], [0, plus, 0, 0], [0, 0, plus, 0], [0, 0, 0, minus]]


class CAN(StdGate):
    """Canonical gate exp(-i pi/2 (tx XX + ty YY + tz ZZ)), composed as ZZ @ YY @ XX like the reference."""
    ARITY = 2
    PARAMS = ('tx', 'ty', 'tz')

    @staticmethod
    def operator(tx, ty, tz):
        return (ZZ(tz) @ (YY(ty) @ XX(tx))).tensor

    def inverse(self):
        tx, ty, tz = self.params.values()
        return CAN(-tx, -ty, -tz, *self.qubits)

    def __pow__(self, t):
        tx, ty, tz = self.params.values()
        return CAN(tx * t, ty * t, tz * t, *self.qubits)


class CANONICAL(CAN):
    """Backwards-compatible alias."""


class EXCH(_HalfTurns2):
    @staticmethod
    def operator(t):
        return CAN(t, t, t).tensor


# ---------------------------------------------------------------------------------------------------------
# three-qubit gates
# ---------------------------------------------------------------------------------------------------------

class CCNOT(StdGate):
    ARITY = 3
    HERMITIAN = True

    @staticmethod
    def operator():
        return _perm_matrix(8, swaps=[(6, 7)])


class CSWAP(StdGate):
    ARITY = 3
    HERMITIAN = True

    @staticmethod
    def operator():
        return _perm_matrix(8, swaps=[(5, 6)])


GATESET = frozenset([I, X, Y, Z, H, S, T, PHASE, RX, RY, RZ, CZ, CNOT, SWAP, ISWAP, CPHASE00, CPHASE01, CPHASE10,
                     CPHASE, PSWAP, CCNOT, CSWAP, PISWAP, RN, TX, TY, TZ, TH, ZYZ, CAN, XX, YY, ZZ, EXCH, S_H,
                     T_H])

STDGATES = {cls.__name__: cls for cls in GATESET}

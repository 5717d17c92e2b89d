"""Measurement, reset, classical-control elements a Circuit may hold, and the classical-memory operations of
hybrid programs (quantumflow/stdops.py:170-341: they only touch `State.memory`, never the amplitudes).

Behavioural contract: quantumflow/stdops.py:38-130 (Measure, Reset, Barrier) and :133-170 (If). `Measure.run` is
the reference's project / roll / renormalise (stdops.py:53-65) fused into two kernels: one marginal reduction
(the probability, read back because the RNG roll happens on the host with `np.random.random()`, exactly one
draw after p0 is known) and one collapse+rescale sweep.
"""
import operator
from typing import Any, Callable, Union

import numpy as np

from . import backend as bk
from . import engine
from .cbits import Addr
from .ops import Channel, Gate, Operation
from .qubits import Qubit, Qubits, QubitVector, asarray
from .states import Density, State

__all__ = ['Measure', 'Reset', 'Barrier', 'If',
           'Neg', 'Not',
           'And', 'Ior', 'Or', 'Xor',
           'Add', 'Mul', 'Sub', 'Div',
           'Move', 'Exchange',
           'EQ', 'LT', 'GT', 'LE', 'GE', 'NE']


class Measure(Operation):
    """Measure one qubit in the computational basis; optionally record the outcome in a classical bit."""

    def __init__(self, qubit: Qubit, cbit: Addr = None) -> None:
        self.qubit = qubit
        self.cbit = cbit

    @property
    def qubits(self) -> Qubits:
        return [self.qubit]

    def quil(self) -> str:
        if self.cbit is not None:
            return '{} {} {}'.format(self.name.upper(), self.qubit, self.cbit)
        return '{} {}'.format(self.name.upper(), self.qubit)

    def run(self, ket: State) -> State:
        count = ket.qubit_nb
        bit = count - 1 - ket.qubits.index(self.qubit)
        p0, p1 = (float(v) for v in asarray(engine.marginal(ket.tensor, bit)))
        outcome = 0 if np.random.random() < p0 else 1
        weight = p0 if outcome == 0 else p1
        tensor = engine.collapse(ket.tensor, bit, outcome, 1.0 / np.sqrt(weight))
        ket = State(tensor, ket.qubits, ket.memory)
        if self.cbit is not None:
            ket = ket.update({self.cbit: outcome})
        return ket

    def evolve(self, rho: Density) -> Density:
        count = rho.qubit_nb
        where = rho.qubits.index(self.qubit)
        ket_bit, bra_bit = 2 * count - 1 - where, count - 1 - where

        def project(value: int) -> Density:
            tensor = engine.collapse(rho.tensor, ket_bit, value, 1.0)
            tensor = engine.collapse(tensor, bra_bit, value, 1.0, inplace=True)
            return Density(tensor, rho.qubits, rho.memory)

        zero = project(0)
        # the reference rolls against the Hilbert-Schmidt norm of P0 rho P0 (stdops.py:71, Density.norm)
        prob_zero = float(asarray(zero.norm()))
        if np.random.random() < prob_zero:
            result, outcome = zero.normalize(), 0
        else:
            result, outcome = project(1).normalize(), 1
        if self.cbit is not None:
            result = result.update({self.cbit: outcome})
        return result


class Reset(Operation):
    """Send qubits to |0> whatever their state: apply [[1,1],[0,0]] per qubit, then renormalise."""

    def __init__(self, *qubits: Qubit) -> None:
        self._qubits = tuple(qubits)
        self.vec = QubitVector([[1, 1], [0, 0]], [0], resident=False)

    @property
    def H(self) -> 'Reset':
        return self

    def run(self, ket: State) -> State:
        qubits = self.qubits if self.qubits else ket.qubits
        tensor = ket.tensor
        for q in qubits:
            tensor = bk.tensormul(self.vec.tensor, tensor, [ket.qubits.index(q)])
        return State(tensor, ket.qubits, ket.memory).normalize()

    def evolve(self, rho: Density) -> Density:
        raise TypeError('Not yet implemented')

    def asgate(self) -> Gate:
        raise TypeError('Reset not convertible to Gate')

    def aschannel(self) -> Channel:
        raise TypeError('Reset not convertible to Channel')

    def quil(self) -> str:
        if self.qubits:
            return 'RESET ' + ' '.join(str(q) for q in self.qubits)
        return 'RESET'


class Barrier(Operation):
    """Does nothing to the state; stops the planner from fusing across it."""

    def __init__(self, *qubits: Qubit) -> None:
        self._qubits = qubits

    @property
    def H(self) -> 'Barrier':
        return self

    def run(self, ket: State) -> State:
        return ket

    def evolve(self, rho: Density) -> Density:
        return rho

    def quil(self) -> str:
        return self.name.upper() + ' ' + ' '.join(str(q) for q in self.qubits)


class If(Operation):
    """Apply `elem` when classical bit `condition` equals `value`."""

    def __init__(self, elem: Operation, condition: Addr, value: Any = True) -> None:
        self.element = elem
        self.condition = condition
        self.value = value

    @property
    def qubits(self) -> Qubits:
        return self.element.qubits

    def run(self, ket: State) -> State:
        if ket.memory[self.condition] == self.value:
            ket = self.element.run(ket)
        return ket

    def evolve(self, rho: Density) -> Density:
        if rho.memory[self.condition] == self.value:
            rho = self.element.evolve(rho)
        return rho


# ---------------------------------------------------------------------------------------------------------
# classical memory operations (reference stdops.py:170-341). The amplitude tensor is passed through untouched
# (State.update re-wraps the same device buffer), so run() and evolve() coincide.
# ---------------------------------------------------------------------------------------------------------

class _Classical(Operation):
    """An operation on classical memory only."""

    def _apply(self, memory) -> dict:
        raise NotImplementedError()

    def run(self, ket: State) -> State:
        return ket.update(self._apply(ket.memory))

    def evolve(self, rho: Density) -> Density:
        return rho.update(self._apply(rho.memory))


class _Unary(_Classical):
    _fn: Callable = None

    def __init__(self, target: Addr) -> None:
        self.target = target
        self.addresses = [target]

    def _apply(self, memory) -> dict:
        return {self.target: type(self)._fn(memory[self.target])}

    def quil(self) -> str:
        return '{} {}'.format(self.name, self.target)


class Neg(_Unary):
    """target <- -target"""
    _fn = staticmethod(operator.neg)


class Not(_Unary):
    """target <- int(not target)"""
    _fn = staticmethod(lambda value: int(not value))


class BinaryOP(_Classical):
    """target <- op(target, source); `source` is an address or an immediate number."""
    _fn: Callable = None

    def __init__(self, target: Addr, source: Union[Addr, int, float]) -> None:
        self.target = target
        self.source = source

    def _source(self, memory):
        return memory[self.source] if isinstance(self.source, Addr) else self.source

    def _apply(self, memory) -> dict:
        return {self.target: type(self)._fn(memory[self.target], self._source(memory))}

    def quil(self) -> str:
        return '{} {} {}'.format(self.name, self.target, self.source)


def _binary(name: str, fn: Callable, doc: str) -> type:
    return type(name, (BinaryOP,), {'_fn': staticmethod(fn), '__doc__': doc, '__module__': __name__})


And = _binary('And', operator.and_, 'target <- target & source')
Ior = _binary('Ior', operator.or_, 'target <- target | source')
Or = _binary('Or', operator.or_, 'target <- target | source (deprecated Quil spelling of Ior)')
Xor = _binary('Xor', operator.xor, 'target <- target ^ source')
Add = _binary('Add', operator.add, 'target <- target + source')
Sub = _binary('Sub', operator.sub, 'target <- target - source')
Mul = _binary('Mul', operator.mul, 'target <- target * source')
Div = _binary('Div', operator.truediv, 'target <- target / source')
Move = _binary('Move', lambda target, source: source, 'target <- source')


class Exchange(BinaryOP):
    """Swap two classical cells."""

    def _apply(self, memory) -> dict:
        assert isinstance(self.source, Addr)
        return {self.target: memory[self.source], self.source: memory[self.target]}


class Comparison(_Classical):
    """target <- op(left, right)"""
    _fn: Callable = None

    def __init__(self, target: Addr, left: Addr, right: Addr) -> None:
        self.target = target
        self.left = left
        self.right = right

    def _apply(self, memory) -> dict:
        return {self.target: type(self)._fn(memory[self.left], memory[self.right])}

    def quil(self) -> str:
        return '{} {} {} {}'.format(self.name, self.target, self.left, self.right)


def _comparison(name: str, fn: Callable) -> type:
    return type(name, (Comparison,), {'_fn': staticmethod(fn), '__module__': __name__,
                                      '__doc__': 'target <- (left {} right)'.format(fn.__name__)})


EQ = _comparison('EQ', operator.eq)
GT = _comparison('GT', operator.gt)
GE = _comparison('GE', operator.ge)
LT = _comparison('LT', operator.lt)
LE = _comparison('LE', operator.le)
NE = _comparison('NE', operator.ne)

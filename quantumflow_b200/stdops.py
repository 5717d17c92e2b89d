"""Measurement, reset and the classical-control elements a Circuit may hold.

Behavioural contract: quantumflow/stdops.py:38-130 (Measure, Reset, Barrier) and :133-170 (If). `Measure.run` is
the reference's project / roll / renormalise (stdops.py:53-65) fused into two kernels: one marginal reduction
(the probability, read back because the RNG roll happens on the host with `np.random.random()`, exactly one
draw after p0 is known) and one collapse+rescale sweep.
"""
from typing import Any

import numpy as np

from . import backend as bk
from . import engine
from .cbits import Addr
from .ops import Channel, Gate, Operation
from .qubits import Qubit, Qubits, QubitVector, asarray
from .states import Density, State

__all__ = ['Measure', 'Reset', 'Barrier', 'If']


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

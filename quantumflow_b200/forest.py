"""The QAM-shaped front of `Program.run` (quantumflow/forest/__init__.py:372-447,
`QuantumFlowQVM`): load a program, run it on the device, read classical memory regions and the wavefunction back.

The reference's class derives from `pyquil.api.QAM` and loads pyQuil programs through pyQuil's Quil parser
(forest/__init__.py:216-219, 389-397). pyQuil is not a dependency here and the Quil TEXT parser is outside the hot
path, so `load` takes a `quantumflow_b200.Program` -- what `quil_to_program` produces in the
reference -- and everything behind it is the reference's state machine ('connected' -> 'loaded' -> 'running' ->
'done'), with the state resident in HBM between `run()` and the read-outs. `wavefunction()` returns the flat amplitude
vector in pyQuil's (bit-reversed) order, i.e. `pyquil.Wavefunction.amplitudes` (forest/__init__.py:350-358)."""
from typing import Optional, Sequence

import numpy as np

from .cbits import Register
from .programs import Program
from .states import State

__all__ = ['QuantumFlowQVM']


class QuantumFlowQVM:
    """A Quantum Virtual Machine over `Program.run` (forest/__init__.py:372-447)."""

    def __init__(self) -> None:
        self.program: Optional[Program] = None
        self.status = 'connected'
        self._prog: Optional[Program] = None
        self._ket: Optional[State] = None

    def load(self, binary) -> 'QuantumFlowQVM':
        """Load a program and initialise the machine into a fresh state (forest/__init__.py:383-397)."""
        assert self.status in ['connected', 'done']
        if isinstance(binary, str):
            raise NotImplementedError('Quil text needs pyQuil\'s parser (forest.quil_to_program); pass a Program')
        if not isinstance(binary, Program):
            raise TypeError('load() takes a quantumflow_b200.Program')
        self._prog = binary
        self.program = binary
        self._ket = None
        self.status = 'loaded'
        return self

    def write_memory(self, *, region_name: str, offset: int = 0, value: int = None) -> 'QuantumFlowQVM':
        raise NotImplementedError()            # as in the reference (forest/__init__.py:399-402)

    def run(self) -> 'QuantumFlowQVM':
        """Run a previously loaded program (forest/__init__.py:404-413); the status stays 'running' until wait(),
        which is what pyQuil's QuantumComputer expects."""
        assert self.status in ['loaded']
        self.status = 'running'
        self._ket = self._prog.run()
        return self

    def wait(self) -> 'QuantumFlowQVM':
        assert self.status == 'running'
        self.status = 'done'
        return self

    def read_from_memory_region(self, *, region_name: str, offsets: Sequence[int] = None) -> Sequence[int]:
        """Values of the classical register `region_name`, in the order the addresses entered the state's memory
        (forest/__init__.py:418-435)."""
        assert self.status == 'done'
        if offsets is not None:
            raise NotImplementedError('Offsets not yet supported')
        reg = Register(region_name)
        assert self._ket is not None
        return [value for addr, value in self._ket.memory.items() if getattr(addr, 'register', None) == reg]

    def wavefunction(self) -> np.ndarray:
        """Amplitudes of a completed program in pyQuil's order (forest/__init__.py:437-445)."""
        assert self.status == 'done'
        assert self._ket is not None
        from .stateio import state_to_wavefunction_amplitudes
        return state_to_wavefunction_amplitudes(self._ket)

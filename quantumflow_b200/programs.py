"""Hybrid quantum-classical programs: the step above `Circuit.run`.

Behavioural contract: quantumflow/programs.py:55-563 -- a `Program` is a list of instructions with Quil control
flow (labels, jumps on classical bits, halt), gate calls by name, declarations and measurements; its interpreter
state (program counter, jump targets, gate table) lives in the classical memory of the State under the private
register `_prog_state_`, exactly as in the reference, so `ket.memory[PC]` / `ket.memory[TARGETS]` read the same.

What is different underneath: the reference interprets one instruction at a time, one `np.einsum` sweep per gate
call (programs.py:152-170, 429-439). Here the amplitudes stay in HBM across the Python control flow, and every
*basic block* -- a maximal run of gate calls between two control-flow / measurement instructions; jumps can only
land on labels, so such a run is always entered at its top -- is handed to the sweep planner as one `Circuit`
(`Circuit.run` / `Circuit.evolve`: many gates per pass over the state, plan cached on the block). A loop body of
40 gates costs a handful of sweeps per iteration instead of 40.
"""
from abc import ABC
from typing import Dict, Generator, List, Optional, Tuple

from .cbits import Addr, Register
from .circuits import Circuit
from .ops import Gate
from .qubits import Qubits
from .states import Density, State, zero_state
from .stdgates import STDGATES

__all__ = ['Instruction', 'Program', 'DefCircuit', 'Wait', 'Nop', 'Halt', 'Label', 'Jump', 'JumpWhen',
           'JumpUnless', 'Pragma', 'Include', 'Call', 'Declare', 'Load', 'Store']
# plus `Parameter` (sympy's Symbol, programs.py:20 of the reference), resolved lazily below and by the package

# private register that holds the interpreter state (programs.py:43-50)
_prog_state_ = Register('_prog_state_')
PC = _prog_state_['pc']
NAMEDGATES = _prog_state_['namedgates']
TARGETS = _prog_state_['targets']
WAIT = _prog_state_['wait']

HALTED = -1             # program counter of a finished program

# gate calls are fused from this many in a row (below that the per-gate kernels are as good)
BLOCK_MIN_CALLS = 2


def __getattr__(name: str):
    # `Parameter` is sympy's Symbol in the reference (programs.py:20); sympy is imported on first use only
    if name == 'Parameter':
        from sympy import Symbol
        return Symbol
    raise AttributeError(name)


class Instruction(ABC):
    """A program instruction: control flow, declarations, gate calls."""

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

    def quil(self) -> str:
        return self.name

    def __str__(self) -> str:
        return self.quil()

    def run(self, ket: State) -> State:
        raise NotImplementedError()

    def evolve(self, rho: Density) -> Density:
        # purely classical instructions act the same on both
        res = self.run(rho)
        assert isinstance(res, Density)
        return res


class Program(Instruction):
    """A program for a hybrid quantum computer, following the Quil instruction set."""

    def __init__(self, instructions: List[Instruction] = None, name: str = None, params: dict = None) -> None:
        self.instructions = [] if instructions is None else instructions
        self._blocks: Dict[Tuple, Dict[int, Tuple[int, Circuit]]] = {}

    def quil(self) -> str:
        return '\n'.join([str(i) for i in self.instructions] + [''])

    @property
    def qubits(self) -> Qubits:
        return sorted({q for instr in self.instructions for q in instr.qubits})

    def __iadd__(self, other: Instruction) -> 'Program':
        self.instructions.append(other)
        return self

    def __len__(self) -> int:
        return len(self.instructions)

    def __getitem__(self, key: int) -> Instruction:
        return self.instructions[key]

    def __iter__(self) -> Generator[Instruction, None, None]:
        for inst in self.instructions:
            yield inst

    def _initilize(self, state: State) -> State:      # (sic) the reference's spelling, programs.py:138
        targets = {instr.target: pc for pc, instr in enumerate(self) if isinstance(instr, Label)}
        return state.update({PC: 0, TARGETS: targets, NAMEDGATES: STDGATES.copy()})

    # ---- basic blocks ------------------------------------------------------------------------------
    def _basic_blocks(self) -> Dict[int, Tuple[int, Circuit]]:
        """start pc -> (end pc, Circuit) for every maximal run of >= BLOCK_MIN_CALLS gate calls. Cached on the
        identities of the instructions, so a program that is extended after a run is analysed again."""
        # keyed on the instructions AND on what a Call resolves to: changing Call.params / gatename / qubits after a
        # run must not replay the stale fused circuit (the reference re-resolves every instruction on each run)
        key = tuple((id(i), i.gatename, tuple(map(_hashable, i.params)), tuple(i.qubits)) if isinstance(i, Call)
                    else id(i) for i in self.instructions)
        hit = self._blocks.get(key)
        if hit is not None:
            return hit
        blocks: Dict[int, Tuple[int, Circuit]] = {}
        pc, n = 0, len(self.instructions)
        while pc < n:
            gates = []
            end = pc
            while end < n:
                gate = _as_gate(self.instructions[end])
                if gate is None:
                    break
                gates.append(gate)
                end += 1
            if len(gates) >= BLOCK_MIN_CALLS:
                blocks[pc] = (end, Circuit(gates))
            pc = max(end, pc + 1)
        self._blocks = {key: blocks}
        return blocks

    def _interpret(self, state: State, foreign_ptr: Optional[int], step, block_step) -> State:
        """The fetch / execute loop. `foreign_ptr`: device address of the CALLER's amplitude buffer (None when the
        program created the state): a block may update the buffer in place only once the state is an intermediate
        of this run, i.e. lives somewhere else."""
        state = self._initilize(state)
        blocks = self._basic_blocks()
        pc = 0
        while 0 <= pc < len(self.instructions):
            block = blocks.get(pc)
            if block is not None:
                end, circuit = block
                owned = foreign_ptr is None or state.tensor.data_ptr() != foreign_ptr
                state = _keep_memory(block_step(circuit, state, owned), state).update({PC: end})
            else:
                state = step(self.instructions[pc], state.update({PC: pc + 1}))
            pc = state.memory[PC]
        return state

    def run(self, ket: State = None) -> State:
        """Run the program on `ket` (default: |0...0> on the program's qubits, empty classical memory)."""
        foreign = None if ket is None else ket.tensor.data_ptr()
        if ket is None:
            ket = zero_state(self.qubits)
        return self._interpret(ket, foreign, lambda instr, state: instr.run(state),
                               lambda circuit, state, owned: circuit.run(state, _owned=owned))

    def evolve(self, rho: Density = None) -> Density:
        foreign = None if rho is None else rho.tensor.data_ptr()
        if rho is None:
            rho = zero_state(self.qubits).asdensity()
        res = self._interpret(rho, foreign, lambda instr, state: instr.evolve(state),
                              lambda circuit, state, owned: circuit.evolve(state, _owned=owned))
        assert isinstance(res, Density)
        return res


def _keep_memory(new: State, old: State) -> State:
    """A block of unitary gates never touches classical memory; carry the interpreter's over whatever the circuit
    executor returned."""
    return type(new)(new.tensor, new.qubits, old.memory)


def _hashable(value):
    try:
        hash(value)
        return value
    except TypeError:
        return id(value)


def _as_gate(instr) -> Optional[Gate]:
    """The Gate a block may fuse for this instruction, or None (control flow, measurement, unknown gate name --
    the latter must raise when it is *executed*, like the reference, not when the program is analysed)."""
    if isinstance(instr, Call):
        if instr.gatename not in STDGATES:
            return None
        try:
            return instr.gate(STDGATES)
        except Exception:        # wrong arity / symbolic parameter: let the per-instruction path raise
            return None
    if isinstance(instr, Gate) and not getattr(instr.tensor, 'requires_grad', False):
        return instr
    return None


class DefCircuit(Program):
    """A named, parameterised sub-program."""

    def __init__(self, name: str, params: Dict[str, float], qubits: Qubits = None,
                 instructions: List[Instruction] = None) -> None:
        super().__init__(instructions)
        self.progname = name
        self.params = params
        self._qubits = [] if qubits is None else qubits

    @property
    def qubits(self) -> Qubits:
        return self._qubits

    def quil(self) -> str:
        fparams = '(' + ','.join(map(str, self.params)) + ')' if self.params else ''
        fqubits = ' ' + ' '.join(map(str, self.qubits)) if self.qubits else ''
        lines = ['{} {}{}{}:'.format(self.name, self.progname, fparams, fqubits)]
        lines += ['    ' + str(instr) for instr in self.instructions]
        return '\n'.join(lines) + '\n'


class _NoEffect(Instruction):
    """Instructions the interpreter records but does not act upon."""

    def run(self, ket: State) -> State:
        return ket


class Wait(_NoEffect):
    """Hand control back to the caller (recorded, no effect)."""


class Nop(_NoEffect):
    """No operation."""


class Halt(Instruction):
    """Stop the program: the program counter becomes HALTED."""

    def run(self, ket: State) -> State:
        return ket.update({PC: HALTED})


class _MemoryTransfer(Instruction):
    """LOAD / STORE between classical memory regions: printable, not executable (as in the reference)."""

    def __init__(self, target, left, right) -> None:
        self.target, self.left, self.right = target, left, right

    def quil(self) -> str:
        return ' '.join(str(part) for part in (self.name, self.target, self.left, self.right))

    def run(self, ket: State) -> State:
        raise NotImplementedError()


class Load(_MemoryTransfer):
    """LOAD target left right"""


class Store(_MemoryTransfer):
    """STORE target left right"""


class _Targeted(Instruction):
    """Instructions that name a label."""

    def __init__(self, target: str) -> None:
        self.target = target

    def quil(self) -> str:
        return '{} @{}'.format(self.name, self.target)


class Label(_Targeted, _NoEffect):
    """A jump target."""


class Jump(_Targeted):
    """Unconditional jump."""

    def run(self, ket: State) -> State:
        return ket.update({PC: ket.memory[TARGETS][self.target]})


class _ConditionalJump(_Targeted):
    _jump_when: bool = True

    def __init__(self, target: str, condition: Addr) -> None:
        super().__init__(target)
        self.condition = condition

    def quil(self) -> str:
        return '{} {}'.format(super().quil(), self.condition)

    def run(self, ket: State) -> State:
        memory = ket.memory
        if bool(memory[self.condition]) == self._jump_when:
            return ket.update({PC: memory[TARGETS][self.target]})
        return ket


class JumpWhen(_ConditionalJump):
    """Jump if the classical bit is one."""
    _jump_when = True

    @property
    def name(self) -> str:
        return 'JUMP-WHEN'


class JumpUnless(_ConditionalJump):
    """Jump if the classical bit is zero."""
    _jump_when = False

    @property
    def name(self) -> str:
        return 'JUMP-UNLESS'


class Pragma(_NoEffect):
    """PRAGMA <command> <arg>* "<freeform>"? -- recorded, no effect."""

    def __init__(self, command: str, args: List[float] = None, freeform: str = None) -> None:
        self.command, self.args, self.freeform = command, args, freeform

    def quil(self) -> str:
        parts = ['PRAGMA {}'.format(self.command)]
        if self.args:
            parts.extend(str(a) for a in self.args)
        if self.freeform:
            parts.append('"{}"'.format(self.freeform))
        return ' '.join(parts)


class Include(Instruction):
    """INCLUDE "file" -- recorded, not acted upon."""

    def __init__(self, filename: str, program: Program = None) -> None:
        self.filename, self.program = filename, program

    def quil(self) -> str:
        return '{} "{}"'.format(self.name, self.filename)

    def run(self, ket: State) -> State:
        raise NotImplementedError()


class Call(Instruction):
    """Apply a named gate."""

    def __init__(self, name: str, params: list, qubits: Qubits) -> None:
        self.gatename = name
        self.params = params
        self._qubits = qubits

    def quil(self) -> str:
        fqubits = ' ' + ' '.join(str(q) for q in self.qubits) if self.qubits else ''
        fparams = '(' + ', '.join(str(p) for p in self.params) + ')' if self.params else ''
        return '{}{}{}'.format(self.gatename, fparams, fqubits)

    def gate(self, namedgates: dict) -> Gate:
        if self.gatename not in namedgates:
            raise RuntimeError('Unknown named gate')
        return namedgates[self.gatename](*self.params).relabel(self.qubits)

    def run(self, ket: State) -> State:
        return self.gate(ket.memory[NAMEDGATES]).run(ket)

    def evolve(self, rho: Density) -> Density:
        return self.gate(rho.memory[NAMEDGATES]).evolve(rho)


class Declare(Instruction):
    """DECLARE name type [size]: zero-initialised classical memory."""

    def __init__(self, memory_name: str, memory_type: str, memory_size: int, shared_region: str = None,
                 offsets: List[Tuple[int, str]] = None) -> None:
        self.memory_name = memory_name
        self.memory_type = memory_type
        self.memory_size = memory_size
        self.shared_region = shared_region
        self.offsets = offsets

    def quil(self) -> str:
        parts = ['DECLARE', self.memory_name, self.memory_type]
        if self.memory_size != 1:
            parts.append('[{}]'.format(self.memory_size))
        if self.shared_region is not None:
            parts += ['SHARING', self.shared_region]
        return ' '.join(parts)

    def run(self, ket: State) -> State:
        reg = Register(self.memory_name, self.memory_type)
        return ket.update({reg[idx]: 0 for idx in range(self.memory_size)})

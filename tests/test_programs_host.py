"""Host logic of the program interpreter (quantumflow_b200/programs.py) and of the classical-memory operations
(stdops.py) on a stand-in state: control flow never touches the amplitudes, so it is checked without a GPU.
The cases follow the reference's tests/test_programs.py:40-104 and tests/test_stdops.py (classical ops)."""
from collections import defaultdict

import pytest

import quantumflow_b200 as qf
from quantumflow_b200.programs import HALTED, PC, TARGETS


class _Tensor:
    def data_ptr(self):
        return 1234


class FakeState:
    """memory / update / tensor / qubits: what classical instructions and the interpreter loop use."""

    def __init__(self, memory=None):
        self._memory = dict(memory or {})
        self.tensor = _Tensor()
        self.qubits = ()

    @property
    def memory(self):
        return defaultdict(int, self._memory)

    def update(self, memory):
        merged = self.memory
        merged.update(memory)
        return FakeState(merged)


def run(prog, state=None):
    return prog._interpret(state or FakeState(), None, lambda instr, st: instr.run(st),
                           lambda circuit, st, owned: pytest.fail('no gate blocks expected'))


def test_labels_are_compiled_into_targets():
    prog = qf.Program([qf.Label('Here'), qf.Nop(), qf.Label('There')])
    ket = run(prog)
    assert ket.memory[TARGETS] == {'Here': 0, 'There': 2}
    assert ket.memory[PC] == 3


def test_jump_jumpwhen_jumpunless():
    ro = qf.Register()
    prog = qf.Program()
    prog += qf.Move(ro[0], 0)
    prog += qf.Jump('There')
    prog += qf.Not(ro[0])
    prog += qf.Label('There')
    prog += qf.Not(ro[0])
    assert run(prog).memory[ro[0]] == 1
    prog += qf.JumpWhen('There', ro[0])
    assert run(prog).memory[ro[0]] == 0
    prog += qf.Not(ro[0])
    prog += qf.JumpUnless('There', ro[0])
    assert run(prog).memory[ro[0]] == 1


def test_wait_and_halt():
    ro = qf.Register()
    prog = qf.Program([qf.Move(ro[0], 0), qf.Wait(), qf.Not(ro[0]), qf.Wait(), qf.Not(ro[0]), qf.Wait(),
                       qf.Not(ro[0])])
    assert run(prog).memory[ro[0]] == 1
    prog = qf.Program([qf.Move(ro[0], 0), qf.Halt(), qf.Not(ro[0])])
    ket = run(prog)
    assert ket.memory[PC] == HALTED and ket.memory[ro[0]] == 0


def test_declare_and_tuple_addresses():
    prog = qf.Program([qf.Declare('ro', 'BIT', 3), qf.Move(('a', 0), 7)])
    ket = run(prog)
    ro = qf.Register('ro', 'BIT')
    assert all(ro[i] in ket._memory and ket.memory[ro[i]] == 0 for i in range(3))
    assert ket.memory[('a', 0)] == 7


def test_classical_operations():
    ro = qf.Register()
    mem = {ro[0]: 1, ro[1]: 0, ro[2]: 6, ro[3]: 4}
    cases = [(qf.Neg(ro[2]), ro[2], -6), (qf.Not(ro[0]), ro[0], 0), (qf.Not(ro[1]), ro[1], 1),
             (qf.And(ro[0], ro[1]), ro[0], 0), (qf.Ior(ro[0], ro[1]), ro[0], 1), (qf.Or(ro[1], 1), ro[1], 1),
             (qf.Xor(ro[0], 1), ro[0], 0), (qf.Add(ro[2], ro[3]), ro[2], 10), (qf.Sub(ro[2], 1), ro[2], 5),
             (qf.Mul(ro[2], ro[3]), ro[2], 24), (qf.Div(ro[2], ro[3]), ro[2], 1.5), (qf.Move(ro[1], ro[2]), ro[1], 6),
             (qf.EQ(ro[1], ro[2], ro[3]), ro[1], False), (qf.GT(ro[1], ro[2], ro[3]), ro[1], True),
             (qf.GE(ro[1], ro[2], ro[2]), ro[1], True), (qf.LT(ro[1], ro[2], ro[3]), ro[1], False),
             (qf.LE(ro[1], ro[3], ro[2]), ro[1], True), (qf.NE(ro[1], ro[2], ro[3]), ro[1], True)]
    for op, addr, want in cases:
        assert op.run(FakeState(mem)).memory[addr] == want, op.quil()
        assert op.evolve(FakeState(mem)).memory[addr] == want
    swapped = qf.Exchange(ro[2], ro[3]).run(FakeState(mem)).memory
    assert swapped[ro[2]] == 4 and swapped[ro[3]] == 6


def test_quil_text():
    ro = qf.Register()
    assert qf.Move(ro[0], 1).quil() == 'MOVE ro[0] 1' and qf.Not(ro[1]).quil() == 'NOT ro[1]'
    assert qf.EQ(ro[0], ro[1], ro[2]).quil() == 'EQ ro[0] ro[1] ro[2]'
    assert str(qf.Program([qf.Call('BELL', params=[], qubits=[])])) == 'BELL\n'
    assert qf.Call('RX', [0.5], [3]).quil() == 'RX(0.5) 3' and qf.Call('CNOT', [], [0, 1]).quil() == 'CNOT 0 1'
    assert qf.Include('somefile.quil', qf.Program()).quil() == 'INCLUDE "somefile.quil"'
    assert qf.Label('x').quil() == 'LABEL @x' and qf.Jump('x').quil() == 'JUMP @x'
    assert qf.JumpWhen('x', ro[0]).quil() == 'JUMP-WHEN @x ro[0]'
    assert qf.JumpUnless('x', ro[0]).quil() == 'JUMP-UNLESS @x ro[0]'
    assert qf.Pragma('gate_time', ['H', 1.5], 'freeform').quil() == 'PRAGMA gate_time H 1.5 "freeform"'
    assert qf.Declare('ro', 'BIT', 4).quil() == 'DECLARE ro BIT [4]' and qf.Declare('x', 'REAL', 1).quil() == 'DECLARE x REAL'
    assert qf.Declare('x', 'OCTET', 2, 'ro').quil() == 'DECLARE x OCTET [2] SHARING ro'
    assert qf.Halt().quil() == 'HALT' and qf.Nop().qubits == () and qf.Nop().qubit_nb == 0
    circ = qf.DefCircuit('bell', {}, instructions=[qf.Call('H', [], [0]), qf.Call('CNOT', [], [0, 1])])
    assert circ.quil() == 'DEFCIRCUIT bell:\n    H 0\n    CNOT 0 1\n'
    assert qf.DefCircuit('rot', {'%theta': 0.5}).quil() == 'DEFCIRCUIT rot(%theta):\n'


def test_basic_blocks_split_at_control_flow_and_unknown_gates():
    ro = qf.Register()
    prog = qf.Program([qf.Call('H', [], [0]), qf.Call('CNOT', [], [0, 1]), qf.X(1),          # block 0..3
                       qf.Label('loop'), qf.Call('RX', [0.3], [0]),                           # single call: no block
                       qf.Measure(0, ro[0]), qf.Call('Y', [], [1]), qf.Call('NOSUCH', [], [0]), qf.Call('Z', [], [1]),
                       qf.Call('T', [], [0]), qf.Call('S', [], [1]), qf.JumpUnless('loop', ro[0])])
    blocks = prog._basic_blocks()
    assert sorted((start, end) for start, (end, _c) in blocks.items()) == [(0, 3), (8, 11)]
    assert [type(g).__name__ for g in blocks[0][1].elements] == ['H', 'CNOT', 'X']
    assert prog._basic_blocks() is blocks                      # cached
    prog += qf.Nop()
    assert prog._basic_blocks() is not blocks                  # a changed program is analysed again
    assert prog.qubits == [0, 1]
    with pytest.raises(RuntimeError):
        qf.Call('NOSUCH', [], [0]).gate({})


def test_classical_operations_inside_a_circuit():
    # reference tests/test_stdops.py:71-108 (test_logics), :111-130, :155-187; the circuit runs on a stand-in
    # state: classical elements never touch the amplitudes
    FakeState.qubit_nb = 0
    c = qf.Register('c')
    circ = qf.Circuit([qf.Move(c[0], 0), qf.Move(c[1], 1), qf.And(c[0], c[1])])
    ket = circ.run(FakeState())
    assert ket.memory == {c[0]: 0, c[1]: 1}
    circ += qf.Not(c[1])
    circ += qf.And(c[0], c[1])
    ket = circ.run(ket)
    assert ket.memory == {c[0]: 0, c[1]: 0}
    ket = qf.Circuit([qf.Move(c[0], 0), qf.Move(c[1], 1), qf.Ior(c[0], c[1])]).run(FakeState())
    assert ket.memory == {c[0]: 1, c[1]: 1}
    circ = qf.Circuit([qf.Move(c[0], 1), qf.Move(c[1], 1), qf.Xor(c[0], c[1])])
    ket = circ.run(FakeState())
    assert ket.memory == {c[0]: 0, c[1]: 1}
    circ += qf.Exchange(c[0], c[1])
    ket = circ.run(ket)
    assert ket.memory == {c[0]: 1, c[1]: 0}
    circ += qf.Move(c[0], c[1])
    ket = circ.run(ket)
    assert ket.memory == {c[0]: 0, c[1]: 0}
    assert str(qf.Neg(c[10])) == 'NEG c[10]'
    ro = qf.Register()
    assert run(qf.Program([qf.Move(ro[0], 1), qf.Move(ro[1], 2), qf.Add(ro[0], ro[1]), qf.Add(ro[0], 4)])).memory[ro[0]] == 7
    assert run(qf.Program([qf.Move(ro[0], 1), qf.Move(ro[1], 2), qf.Mul(ro[0], ro[1]), qf.Mul(ro[0], 4)])).memory[ro[0]] == 8
    assert run(qf.Program([qf.Move(ro[0], 4), qf.Move(ro[1], 1), qf.Div(ro[0], ro[1]), qf.Div(ro[0], 2)])).memory[ro[0]] == 2
    assert run(qf.Program([qf.Move(ro[0], 1), qf.Move(ro[1], 2), qf.Sub(ro[0], ro[1]), qf.Sub(ro[0], 4),
                           qf.Neg(ro[0])])).memory[ro[0]] == 5
    ket = run(qf.Program([qf.Move(ro[0], 1), qf.Move(ro[1], 2), qf.EQ(('eq', 0), ro[0], ro[1]),
                          qf.GT(('gt', 0), ro[0], ro[1]), qf.GE(('ge', 0), ro[0], ro[1]),
                          qf.LT(('lt', 0), ro[0], ro[1]), qf.LE(('le', 0), ro[0], ro[1])]))
    assert [ket.memory[(k, 0)] for k in ('eq', 'gt', 'ge', 'lt', 'le')] == [0, 0, 0, 1, 1]


def test_registers_and_addresses():
    # reference tests/test_cbits.py:12-53
    ro = qf.Register()
    assert ro.name == 'ro' and str(ro) == "Register('ro', 'BIT')"
    assert qf.Register() == qf.Register('ro') and qf.Register('a') < qf.Register('b')
    assert qf.Register('a') != qf.Register('b') and qf.Register('c') != 'foobar'
    with pytest.raises(TypeError):
        qf.Register('c') < 'foobar'
    c0 = qf.Register('c')[0]
    assert c0.register.name == 'c' and c0.key == 0 and c0.register.dtype == 'BIT'
    assert str(c0) == 'c[0]' and repr(c0) == "Register('c', 'BIT')[0]"
    assert {c0: '1234'}[qf.Register('c')[0]] == '1234'
    assert qf.Register('c')[0] != qf.Register('c')[1] and qf.Register('d')[0] != qf.Register('c')[0]
    assert qf.Register('c')[0] != 'foobar' and qf.Register('c')[0] < qf.Register('c')[1]
    with pytest.raises(TypeError):
        qf.Register('c')[0] < 'foobar'


def test_qvm_state_machine_and_memory_readout():
    """forest.QuantumFlowQVM (quantumflow/forest/__init__.py:372-447; the reference's tests/test_forest.py:119-193) on a
    program of classical instructions only, run on the stand-in state: status transitions, read-out of a memory
    region in the order its addresses entered the memory, the NotImplemented paths."""
    from quantumflow_b200 import forest
    prog = qf.Program([qf.Declare('ro', 'BIT', 2), qf.Move(qf.Register('ro')[1], 1), qf.Move(qf.Register('b')[0], 1)])
    prog.run = lambda ket=None: run(prog)           # no amplitudes involved: interpret on the stand-in state
    qvm = forest.QuantumFlowQVM()
    assert qvm.status == 'connected'
    with pytest.raises(NotImplementedError):
        qvm.load('H 0')                              # Quil text needs pyQuil's parser
    with pytest.raises(TypeError):
        qvm.load(42)
    assert qvm.load(prog) is qvm and qvm.status == 'loaded' and qvm.program is prog
    with pytest.raises(AssertionError):
        qvm.wait()                                   # nothing is running yet
    with pytest.raises(NotImplementedError):
        qvm.write_memory(region_name='ro')
    assert qvm.run() is qvm and qvm.status == 'running'
    with pytest.raises(AssertionError):
        qvm.read_from_memory_region(region_name='ro')
    assert qvm.wait().status == 'done'
    assert qvm.read_from_memory_region(region_name='ro') == [0, 1]
    assert qvm.read_from_memory_region(region_name='missing') == []
    with pytest.raises(NotImplementedError):
        qvm.read_from_memory_region(region_name='ro', offsets=[1])
    # a finished machine takes the next program
    assert qvm.load(prog).run().wait().read_from_memory_region(region_name='ro') == [0, 1]

"""GPU parity of hybrid programs (quantumflow_b200/programs.py) against the reference's own interpreter: the
fixtures of tests/golden/make_golden_programs.py (amplitudes, classical memory, program counter and the position
of the shared numpy RNG stream after the run) and the cases of the reference's tests/test_programs.py (cited)."""
import json
import os

import numpy as np
import pytest

import quantumflow_b200 as qf
from quantumflow_b200 import engine
from quantumflow_b200.programs import PC, TARGETS

from conftest import AMP_TOL, GOLDEN

pytestmark = pytest.mark.gpu


def build(spec):
    regs = {}

    def addr(a):
        return regs.setdefault(a[0], qf.Register(a[0]))[a[1]]

    make = {'call': lambda it: qf.Call(it[1], list(it[2]), list(it[3])),
            'move': lambda it: qf.Move(addr(it[1]), it[2]),
            'label': lambda it: qf.Label(it[1]),
            'measure': lambda it: qf.Measure(it[1], addr(it[2])),
            'jump_unless': lambda it: qf.JumpUnless(it[1], addr(it[2])),
            'jump_when': lambda it: qf.JumpWhen(it[1], addr(it[2])),
            'not': lambda it: qf.Not(addr(it[1])),
            'halt': lambda it: qf.Halt()}
    return qf.Program([make[item[0]](item) for item in spec])


@pytest.fixture(scope='module')
def golden_programs():
    with open(os.path.join(GOLDEN, 'programs_meta.json')) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, 'programs.npz'))


@pytest.mark.parametrize('name', ['measure_until', 'repeat_until_success', 'branch'])
def test_program_runs_match_reference(golden_programs, name):
    meta, arrays = golden_programs
    runs = [r for r in meta['runs'] if r['program'] == name]
    assert len(runs) == 4
    prog = build(meta['specs'][name])
    for r in runs:
        np.random.seed(r['seed'])
        ket = prog.run()
        probe = float(np.random.random())
        assert probe == r['rng_probe'], 'the run consumed a different number of RNG draws'
        memory = {'{}[{}]'.format(a.register.name, a.key): int(v) for a, v in ket.memory.items()
                  if hasattr(a, 'register') and a.register.name in ('ro', 'c')}
        assert memory == r['memory'] and ket.memory[PC] == r['pc'] and list(ket.qubits) == r['qubits']
        got = qf.asarray(ket.tensor).reshape(-1)
        assert np.abs(got - arrays[r['key']]).max() < AMP_TOL


def test_gate_blocks_run_as_fused_sweeps(golden_programs):
    meta, _ = golden_programs
    prog = build(meta['specs']['repeat_until_success'])
    ncalls = sum(1 for item in meta['specs']['repeat_until_success'] if item[0] == 'call')
    np.random.seed(0)
    prog.run()                      # plans the blocks
    np.random.seed(0)
    before = engine.launch_count()
    prog.run()
    launches = engine.launch_count() - before
    assert 0 < launches < ncalls / 2, (launches, ncalls)


def test_program_does_not_modify_the_callers_state(golden_programs):
    meta, _ = golden_programs
    prog = build(meta['specs']['repeat_until_success'])
    ket0 = qf.random_state(10)
    snapshot = qf.asarray(ket0.tensor).copy()
    np.random.seed(3)
    out = prog.run(ket0)
    assert np.array_equal(qf.asarray(ket0.tensor), snapshot)
    assert abs(float(qf.asarray(out.norm())) - 1) < 1e-12


def test_reference_program_cases():
    # tests/test_programs.py:15-19, 40-48
    ket = qf.Program().run()
    assert ket.qubits == () and ket.qubit_nb == 0
    qf.Program([qf.Nop()]).evolve()
    ket = qf.Program([qf.Label('Here'), qf.Nop(), qf.Label('There')]).run()
    assert ket.memory[TARGETS] == {'Here': 0, 'There': 2}
    # :106-130 reset
    ro = qf.Register()
    prog = qf.Program([qf.Move(ro[0], 1), qf.Call('X', params=[], qubits=[0]), qf.Reset(), qf.Measure(0, ro[1])])
    ket = prog.run()
    assert ket.qubits == (0,) and ket.memory[ro[0]] == 1 and ket.memory[ro[1]] == 0
    prog = qf.Program([qf.Call('X', params=[], qubits=[0]), qf.Call('X', params=[], qubits=[1]), qf.Reset(0),
                       qf.Measure(0, ('b', 0)), qf.Measure(1, ('b', 1))])
    ket = prog.run()
    assert ket.memory[('b', 0)] == 0 and ket.memory[('b', 1)] == 1
    # :166-190 bell, occupation basis
    ket = qf.Program([qf.Call('H', [], [0]), qf.Call('CNOT', [], [0, 1])]).run()
    assert qf.states_close(ket, qf.ghz_state(2))
    ket = qf.Program([qf.Call('X', [], [0]), qf.Call('X', [], [1]), qf.Call('I', [], [2]), qf.Call('I', [], [3])]).run()
    probs = qf.asarray(ket.probabilities())
    assert ket.qubits == (0, 1, 2, 3) and probs[1, 1, 0, 0] == 1.0 and probs[1, 1, 0, 1] == 0.0
    # :312-316
    with pytest.raises(RuntimeError):
        qf.Program([qf.Call('NOT_A_GATE', [], [0])]).run()


def test_qaoa_program_golden_wavefunction():
    # tests/test_programs.py:193-212
    wf_true = [0.00167784 + 1.00210180e-05 * 1j, 0.50000000 - 4.99997185e-01 * 1j,
               0.50000000 - 4.99997185e-01 * 1j, 0.00167784 + 1.00210180e-05 * 1j]
    calls = [('RY', [np.pi / 2], [0]), ('RX', [np.pi], [0]), ('RY', [np.pi / 2], [1]), ('RX', [np.pi], [1]),
             ('CNOT', [], [0, 1]), ('RX', [-np.pi / 2], [1]), ('RY', [4.71572463191], [1]), ('RX', [np.pi / 2], [1]),
             ('CNOT', [], [0, 1]), ('RX', [-2 * 2.74973750579], [0]), ('RX', [-2 * 2.74973750579], [1])]
    prog = qf.Program([qf.Call(*c) for c in calls])
    assert qf.states_close(prog.run(), qf.State(wf_true))
    rho = prog.evolve()
    assert qf.densities_close(rho, qf.State(wf_true).asdensity())


def test_program_evolve_with_measurement_and_blocks():
    # Program.evolve: gate blocks through Circuit.evolve, Measure.evolve in between. The outcome is random (the
    # density roll differs from the ket roll in the reference too: stdops.py:53-82), so the expectation is built
    # for the outcome that was drawn: project, renormalise, continue.
    ro = qf.Register()
    head = [qf.Call('H', [], [0]), qf.Call('CNOT', [], [0, 1]), qf.Call('RX', [0.3], [2])]
    tail = [qf.Call('RY', [0.7], [1]), qf.Call('CZ', [], [1, 2]), qf.Call('T', [], [2])]
    for seed in range(4):
        np.random.seed(seed)
        rho = qf.Program(head + [qf.Measure(0, ro[0])] + tail).evolve()
        outcome = rho.memory[ro[0]]
        ket = qf.Program(head).run()
        ket = (qf.P1(0) if outcome else qf.P0(0)).run(ket).normalize()
        ket = qf.Program(tail).run(ket)
        assert qf.densities_close(rho, ket.asdensity())
        assert abs(complex(qf.asarray(rho.trace())) - 1) < 1e-12

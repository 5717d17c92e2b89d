#!/usr/bin/env python
"""Golden vectors for hybrid programs (tests/test_gpu_programs.py), produced by the REFERENCE's own interpreter
(quantumflow/programs.py Program.run through tests/golden/refshim.py). Build container only: the reference tree
is not on the GPU box. Writes programs.npz + programs_meta.json next to this script.

The programs are described as plain data (`spec` lists) so that the test rebuilds them with quantumflow_b200's
classes: ('call', name, params, qubits) | ('move', addr, value) | ('label', name) | ('measure', qubit, addr)
| ('jump_unless', label, addr) | ('jump_when', label, addr) | ('not', addr) | ('halt',)
with addr = [register name, key]."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refshim import load_reference                      # noqa: E402
from quantumflow_b200 import workloads                  # noqa: E402  (gate lists only: plain data)


def build(qf, spec):
    regs = {}

    def addr(a):
        reg = regs.setdefault(a[0], qf.Register(a[0]))
        return reg[a[1]]

    prog = qf.Program()
    for item in spec:
        kind = item[0]
        if kind == 'call':
            prog += qf.Call(item[1], list(item[2]), list(item[3]))
        elif kind == 'move':
            prog += qf.Move(addr(item[1]), item[2])
        elif kind == 'label':
            prog += qf.Label(item[1])
        elif kind == 'measure':
            prog += qf.Measure(item[1], addr(item[2]))
        elif kind == 'jump_unless':
            prog += qf.JumpUnless(item[1], addr(item[2]))
        elif kind == 'jump_when':
            prog += qf.JumpWhen(item[1], addr(item[2]))
        elif kind == 'not':
            prog += qf.Not(addr(item[1]))
        elif kind == 'halt':
            prog += qf.Halt()
        else:
            raise ValueError(kind)
    return prog, addr


def layer_calls(n, depth, seed):
    return [('call', name, [float(p) for p in params], [int(q) for q in qubits])
            for name, params, qubits in workloads.wb_gate_list(n, depth, seed)]


def specs():
    out = {}
    # reference tests/test_programs.py:152-163 (measure until the bit reads one)
    out['measure_until'] = [('move', ['c', 2], 1), ('label', 'redo'), ('call', 'X', [], [0]), ('call', 'H', [], [0]),
                            ('measure', 0, ['c', 2]), ('jump_unless', 'redo', ['c', 2])]
    # repeat-until-success around fused gate blocks: 10 qubits, loop body of 2 layers, a tail block of 3 layers
    n = 10
    body = layer_calls(n, 2, 11)
    tail = layer_calls(n, 3, 12)[n:]              # without the leading H layer
    out['repeat_until_success'] = ([('label', 'redo')] + body + [('measure', 3, ['ro', 0]),
                                   ('jump_unless', 'redo', ['ro', 0])] + tail +
                                   [('measure', 7, ['ro', 1]), ('jump_when', 'done', ['ro', 1]),
                                    ('call', 'X', [], [0]), ('label', 'done')] + layer_calls(n, 1, 13)[n:])
    # classical control only decides which block runs
    out['branch'] = ([('call', 'H', [], [0]), ('call', 'CNOT', [], [0, 1]), ('measure', 0, ['ro', 0]),
                      ('jump_when', 'one', ['ro', 0]), ('call', 'RX', [0.4], [2]), ('call', 'CNOT', [], [2, 1]),
                      ('halt',), ('label', 'one'), ('call', 'RY', [1.1], [2]), ('call', 'CZ', [], [2, 0]),
                      ('call', 'T', [], [1])])
    return out


def main():
    qf = load_reference()
    arrays, meta = {}, {'specs': specs(), 'runs': []}
    for name, spec in meta['specs'].items():
        for seed in range(4):
            prog, addr = build(qf, spec)
            np.random.seed(seed)
            ket = prog.run()
            probe = float(np.random.random())           # position of the shared RNG stream after the run
            key = '{}_{}'.format(name, seed)
            arrays[key] = np.asarray(ket.vec.asarray(), dtype=np.complex128).reshape(-1)
            bits = {'{}[{}]'.format(a.register.name, a.key): int(v) for a, v in ket.memory.items()
                    if hasattr(a, 'register') and a.register.name in ('ro', 'c')}
            meta['runs'].append({'program': name, 'seed': seed, 'key': key, 'memory': bits, 'rng_probe': probe,
                                 'pc': int(ket.memory[qf.programs.PC]), 'qubits': [int(q) for q in ket.qubits]})
    np.savez_compressed(os.path.join(HERE, 'programs.npz'), **arrays)
    with open(os.path.join(HERE, 'programs_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('wrote', len(arrays), 'runs')


if __name__ == '__main__':
    main()

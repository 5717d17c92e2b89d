"""Synthetic workload generators (specification: SURVEY.md 8d / Appendix G). They use only the public gate API,
are driven by Python's `random`, and are therefore identical for this package and for the reference.

  wa_circuit  tools/benchmark.py:31-45 verbatim as a Circuit (N x H, then (X, T, CNOT) triples on random pairs)
  wb_circuit  depth-D layered random 1q/2q circuit (configs C1 / C4 / C5 of BASELINE.json)
  wd_circuit  density-matrix workload: RX layer, CNOT matching, depolarizing channel per qubit (config C3)
  wb_gate_count / wb_gate_list: the same circuit as plain (name, params, qubits) tuples, for oracles
"""
import math
import random
from typing import List, Tuple

GateSpec = Tuple[str, Tuple[float, ...], Tuple[int, ...]]


def wa_gate_list(nqubits: int, seed: int, gates: int = 100) -> List[GateSpec]:
    random.seed(seed)
    specs: List[GateSpec] = [('H', (), (n,)) for n in range(nqubits)]
    qubits = list(range(nqubits))
    for _ in range((gates - nqubits) // 3):
        q0, q1 = random.sample(qubits, 2)
        specs += [('X', (), (q0,)), ('T', (), (q1,)), ('CNOT', (), (q0, q1))]
    return specs


def wb_gate_list(nqubits: int, depth: int, seed: int) -> List[GateSpec]:
    rnd = random.Random(seed)
    specs: List[GateSpec] = [('H', (), (q,)) for q in range(nqubits)]
    for _ in range(depth):
        for q in range(nqubits):
            name = rnd.choice(['H', 'X', 'T', 'RX', 'RY', 'RZ'])
            if name in ('RX', 'RY', 'RZ'):
                specs.append((name, (rnd.uniform(0.0, 2 * math.pi),), (q,)))
            else:
                specs.append((name, (), (q,)))
        perm = list(range(nqubits))
        rnd.shuffle(perm)
        for i in range(0, nqubits - 1, 2):
            specs.append((rnd.choice(['CNOT', 'CZ']), (), (perm[i], perm[i + 1])))
    return specs


def wd_gate_list(nqubits: int, depth: int, seed: int, p: float = 0.01) -> List[GateSpec]:
    rnd = random.Random(seed)
    specs: List[GateSpec] = []
    for _ in range(depth):
        for q in range(nqubits):
            specs.append(('RX', (rnd.uniform(0.0, 2 * math.pi),), (q,)))
        perm = list(range(nqubits))
        rnd.shuffle(perm)
        for i in range(0, nqubits - 1, 2):
            specs.append(('CNOT', (), (perm[i], perm[i + 1])))
        for q in range(nqubits):
            specs.append(('DEPOLARIZING', (p,), (q,)))
    return specs


def wb_gate_count(nqubits: int, depth: int) -> int:
    return nqubits + depth * (nqubits + nqubits // 2)


def circuit_from_specs(qf, specs: List[GateSpec], kraus: bool = True):
    """Build a Circuit with any module `qf` that offers the QuantumFlow gate API."""
    circ = qf.Circuit()
    for name, params, qubits in specs:
        if name == 'DEPOLARIZING':
            chan = qf.Depolarizing(params[0], qubits[0])
            circ += chan if kraus else chan.aschannel()
        else:
            circ += getattr(qf, name)(*params, *qubits)
    return circ


def wa_circuit(qf, nqubits: int, seed: int, gates: int = 100):
    return circuit_from_specs(qf, wa_gate_list(nqubits, seed, gates))


def wb_circuit(qf, nqubits: int, depth: int, seed: int):
    return circuit_from_specs(qf, wb_gate_list(nqubits, depth, seed))


def wd_circuit(qf, nqubits: int, depth: int, seed: int, p: float = 0.01, kraus: bool = True):
    return circuit_from_specs(qf, wd_gate_list(nqubits, depth, seed, p), kraus=kraus)

"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (numpy backend) in the build
container. Re-run with:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed here as small .npz/.json files.
Every fixture records the seeds and inputs that produced it, so the tests can rebuild the same inputs with
quantumflow_b200 / the oracle and compare.
"""
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from refshim import load_reference   # noqa: E402

qf = load_reference()
bk = qf.backend

from quantumflow_b200 import workloads   # noqa: E402  (spec generators only: pure `random`, no engine code)

GATE_PARAMS = {
    'I': [], 'X': [], 'Y': [], 'Z': [], 'H': [], 'S': [], 'T': [], 'S_H': [], 'T_H': [],
    'PHASE': [0.7321], 'RX': [1.2345], 'RY': [-2.468], 'RZ': [4.321],
    'RN': [0.9, 0.48, -0.6, 0.64], 'TX': [0.37], 'TY': [1.61], 'TZ': [-0.42], 'TH': [0.83],
    'ZYZ': [0.21, -0.57, 1.3],
    'CZ': [], 'CNOT': [], 'SWAP': [], 'ISWAP': [],
    'CPHASE00': [0.3], 'CPHASE01': [-1.1], 'CPHASE10': [2.2], 'CPHASE': [0.77],
    'PSWAP': [0.55], 'PISWAP': [0.31],
    'CAN': [0.11, 0.23, -0.34], 'XX': [0.45], 'YY': [-0.27], 'ZZ': [0.63], 'EXCH': [0.19],
    'CCNOT': [], 'CSWAP': [],
}


def flat(tensor):
    return np.asarray(bk.evaluate(tensor), dtype=np.complex128).reshape(-1)


def make_stdgates():
    out = {}
    meta = {}
    for name, params in GATE_PARAMS.items():
        gate = qf.STDGATES[name](*params)
        dim = 2 ** gate.qubit_nb
        out[name] = flat(gate.tensor).reshape(dim, dim)
        out[name + '__H'] = flat(gate.H.tensor).reshape(dim, dim)
        try:
            out[name + '__pow'] = flat((gate ** 0.3).tensor).reshape(dim, dim)
        except Exception:           # pragma: no cover
            pass
        meta[name] = params
    out['P0'] = flat(qf.P0().tensor).reshape(2, 2)
    out['P1'] = flat(qf.P1().tensor).reshape(2, 2)
    # composite constructors
    out['control_gate_RX'] = flat(qf.control_gate(5, qf.RX(0.4, 2)).tensor).reshape(4, 4)
    out['conditional_gate'] = flat(qf.conditional_gate(0, qf.X(1), qf.RY(0.3, 1)).tensor).reshape(4, 4)
    out['join_gates'] = flat(qf.join_gates(qf.H(0), qf.CNOT(1, 2)).tensor).reshape(8, 8)
    out['aschannel_RX'] = flat(qf.RX(0.9, 0).aschannel().tensor).reshape(4, 4)
    out['aschannel_CNOT'] = flat(qf.CNOT(0, 1).aschannel().tensor).reshape(16, 16)
    out['depolarizing_superop'] = flat(qf.Depolarizing(0.1, 0).aschannel().tensor).reshape(4, 4)
    out['damping_superop'] = flat(qf.Damping(0.2, 0).aschannel().tensor).reshape(4, 4)
    out['dephasing_superop'] = flat(qf.Dephasing(0.3, 0).aschannel().tensor).reshape(4, 4)
    out['damping_choi'] = np.asarray(bk.evaluate(qf.Damping(0.2, 0).aschannel().choi()))
    np.savez_compressed(os.path.join(HERE, 'stdgates.npz'), **out)
    with open(os.path.join(HERE, 'stdgates_params.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def make_tensormul():
    """bk.tensormul / inner / outer on seeded random inputs (reference numpybk)."""
    rng = np.random.RandomState(1234)
    cases = []
    out = {}
    specs = [(5, 1, [2]), (5, 1, [4]), (6, 2, [4, 1]), (6, 2, [0, 5]), (7, 3, [6, 0, 3]), (7, 3, [1, 2, 3]),
             (8, 4, [7, 2, 5, 0]), (9, 5, [8, 1, 6, 3, 0]), (4, 4, [2, 0, 3, 1]), (10, 6, [9, 0, 4, 7, 2, 5])]
    for n, k, idx in specs:
        gate = rng.normal(size=[2] * (2 * k)) + 1j * rng.normal(size=[2] * (2 * k))
        state = rng.normal(size=[2] * n) + 1j * rng.normal(size=[2] * n)
        res = bk.tensormul(bk.astensor(gate), bk.astensor(state), idx)
        tag = 'tm_{}'.format(len(cases))
        out[tag + '_gate'] = gate
        out[tag + '_state'] = state
        out[tag + '_out'] = np.asarray(bk.evaluate(res))
        cases.append({'tag': tag, 'n': n, 'k': k, 'indices': idx})
    a = rng.normal(size=[2] * 6) + 1j * rng.normal(size=[2] * 6)
    b = rng.normal(size=[2] * 6) + 1j * rng.normal(size=[2] * 6)
    out['inner_a'], out['inner_b'] = a, b
    out['inner_out'] = np.asarray(bk.evaluate(bk.inner(bk.astensor(a), bk.astensor(b))))
    out['outer_out'] = np.asarray(bk.evaluate(bk.outer(bk.astensor(a[0, 0]), bk.astensor(b[1, 1, 0]))))
    rho = rng.normal(size=[2] * 6) + 1j * rng.normal(size=[2] * 6)
    out['productdiag_in'] = rho
    out['productdiag_out'] = np.asarray(bk.evaluate(bk.productdiag(bk.astensor(rho))))
    np.savez_compressed(os.path.join(HERE, 'tensormul.npz'), **out)
    with open(os.path.join(HERE, 'tensormul_cases.json'), 'w') as f:
        json.dump(cases, f, indent=1)


def make_workloads():
    out = {}
    meta = {}
    # W-B N=12 full vector, seeds 0..2
    for seed in (0, 1, 2):
        circ = workloads.wb_circuit(qf, 12, 20, seed)
        out['wb12_seed{}'.format(seed)] = flat(circ.run().tensor)
    # W-B N=16 depth 8 (fits the tile executor with holes), full vector
    circ = workloads.wb_circuit(qf, 16, 8, 3)
    out['wb16_d8_seed3'] = flat(circ.run().tensor)
    # W-A N=12 and 16
    out['wa12_seed0'] = flat(workloads.wa_circuit(qf, 12, 0).run().tensor)
    out['wa16_seed1'] = flat(workloads.wa_circuit(qf, 16, 1).run().tensor)
    # W-B N=20 (config C1): selected amplitudes + moments (the full vector is 16 MiB)
    circ = workloads.wb_circuit(qf, 20, 20, 0)
    ket = flat(circ.run().tensor)
    sel = [0, 1, 524288, 1048575, 123456, 777777, 31337]
    probs = np.abs(ket) ** 2
    meta['wb20_seed0'] = {'indices': sel, 'norm': float(probs.sum()),
                          'mean_index': float((np.arange(ket.size) * probs).sum() / ket.size),
                          'gates': len(circ.elements)}
    out['wb20_seed0_amps'] = ket[sel]
    out['wb20_seed0_stride'] = ket[::4099]          # 256 strided amplitudes
    # density workloads (W-D), Kraus and superoperator paths
    for kraus in (True, False):
        circ = workloads.wd_circuit(qf, 6, 20, 0, kraus=kraus)
        rho = circ.evolve()
        out['wd6_seed0_{}'.format('kraus' if kraus else 'chan')] = flat(rho.tensor).reshape(64, 64)
    circ = workloads.wd_circuit(qf, 8, 4, 1, kraus=True)
    out['wd8_d4_seed1_kraus'] = flat(circ.evolve().tensor).reshape(256, 256)
    # damping variant
    rnd = random.Random(5)
    circ = qf.Circuit()
    for d in range(6):
        for q in range(5):
            circ += qf.RY(rnd.uniform(0, 2 * math.pi), q)
        for q in range(0, 4, 2):
            circ += qf.CNOT(q, q + 1)
        for q in range(5):
            circ += qf.Damping(0.05, q)
    out['damping5_seed5'] = flat(circ.evolve().tensor).reshape(32, 32)
    # library circuits
    out['qft3_of_001'] = flat(qf.qft_circuit([0, 1, 2]).run(qf.Circuit([qf.X(2)]).run(qf.zero_state(3))).tensor)
    out['qft5_random'] = None
    np.random.seed(77)
    ket0 = qf.random_state(5)
    out['qft5_random_in'] = flat(ket0.tensor)
    out['qft5_random'] = flat(qf.qft_circuit([0, 1, 2, 3, 4]).run(ket0).tensor)
    out['ghz12'] = flat(qf.ghz_circuit(range(12)).run().tensor)
    out['ccnot_circuit_gate'] = flat(qf.ccnot_circuit([0, 1, 2]).asgate().tensor).reshape(8, 8)
    add = qf.addition_circuit([0, 1, 2], [3, 4, 5], [6, 7])
    ket = qf.Circuit([qf.X(0), qf.X(2), qf.X(4), qf.X(5)]).run(qf.zero_state(8))
    out['adder3_state'] = flat(add.run(ket).tensor)
    gate = qf.RZ(-4 * np.pi * 0.25, 4)
    pe = qf.phase_estimation_circuit(gate, range(4))
    out['phase_est'] = flat(pe.run().tensor)
    # mixed-arity circuit with 3-qubit gates, PISWAP/CAN etc
    rnd = random.Random(11)
    circ = qf.Circuit()
    names1 = ['H', 'S', 'T', 'X', 'Y', 'Z', 'S_H', 'T_H']
    for d in range(12):
        for q in range(9):
            circ += getattr(qf, rnd.choice(names1))(q)
        a, b, c = rnd.sample(range(9), 3)
        circ += qf.CCNOT(a, b, c)
        a, b, c = rnd.sample(range(9), 3)
        circ += qf.CSWAP(a, b, c)
        a, b = rnd.sample(range(9), 2)
        circ += qf.CAN(rnd.random(), rnd.random(), rnd.random(), a, b)
        a, b = rnd.sample(range(9), 2)
        circ += qf.PISWAP(rnd.random(), a, b)
        a, b = rnd.sample(range(9), 2)
        circ += qf.ISWAP(a, b)
        a, b = rnd.sample(range(9), 2)
        circ += qf.CPHASE(rnd.random(), a, b)
        a, b = rnd.sample(range(9), 2)
        circ += qf.SWAP(a, b)
        circ += qf.TX(rnd.random(), rnd.randrange(9))
        circ += qf.ZYZ(rnd.random(), rnd.random(), rnd.random(), rnd.randrange(9))
    out['mixed9_seed11'] = flat(circ.run().tensor)
    np.savez_compressed(os.path.join(HERE, 'workloads.npz'), **{k: v for k, v in out.items() if v is not None})
    with open(os.path.join(HERE, 'workloads_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def make_sampling():
    """Outcomes of the reference's RNG-consuming calls under fixed seeds (shared-RNG-stream parity)."""
    out = {}
    circ = workloads.wb_circuit(qf, 8, 3, 9)
    ket = circ.run()
    np.random.seed(42)
    out['measure_seq'] = np.asarray([ket.measure() for _ in range(16)])
    np.random.seed(43)
    out['sample_1000'] = ket.sample(1000).reshape(-1)
    # mid-circuit measurement: Measure.run consumes one np.random.random() each
    np.random.seed(44)
    prog = qf.Circuit()
    ro = qf.Register('ro')
    for q in range(6):
        prog += qf.H(q)
    prog += qf.CNOT(0, 1)
    prog += qf.Measure(0, ro[0])
    prog += qf.RX(0.3, 2)
    prog += qf.CNOT(2, 3)
    prog += qf.Measure(3, ro[1])
    prog += qf.Measure(1, ro[2])
    res = prog.run()
    out['midcircuit_state'] = flat(res.tensor)
    out['midcircuit_bits'] = np.asarray([res.memory[ro[i]] for i in range(3)])
    out['midcircuit_next_random'] = np.asarray([np.random.random()])
    # Kraus.run / UnitaryMixture.run stochastic unravelling
    np.random.seed(45)
    ket1 = qf.Circuit([qf.H(0), qf.CNOT(0, 1), qf.RY(0.4, 2)]).run(qf.zero_state(3))
    for q in range(3):
        ket1 = qf.Damping(0.3, q).run(ket1)
        ket1 = qf.Depolarizing(0.5, q).run(ket1)
    out['kraus_run_state'] = flat(ket1.tensor)
    # Measure.evolve on a density
    np.random.seed(46)
    rho = qf.Circuit([qf.H(0), qf.CNOT(0, 1), qf.RX(0.7, 1)]).evolve()
    rho = qf.Measure(1, ro[0]).evolve(rho)
    out['measure_evolve_rho'] = flat(rho.tensor).reshape(4, 4)
    out['measure_evolve_bit'] = np.asarray([rho.memory[ro[0]]])
    # random_state / random_density draw order
    np.random.seed(47)
    out['random_state4'] = flat(qf.random_state(4).tensor)
    out['random_density2'] = flat(qf.random_density(2).tensor).reshape(4, 4)
    # Reset
    np.random.seed(48)
    ket2 = qf.random_state(4)
    out['reset_in'] = flat(ket2.tensor)
    out['reset_out'] = flat(qf.Reset(1, 3).run(ket2).tensor)
    np.savez_compressed(os.path.join(HERE, 'sampling.npz'), **out)


def make_qaoa():
    """QAOA expectation (reference numpy forward) and central finite-difference gradient on the graph of
    tests/test_qaoa_maxcut.py:18 (TF autograd itself is not installable here)."""
    import networkx as nx
    graph = nx.from_edgelist([[0, 1], [1, 2], [1, 3]])
    steps = 5
    beta = np.full(steps, 0.5)
    gamma = np.full(steps, 0.5)
    cuts = qf.graph_cuts(graph)

    def expect(b, g):
        circ = qf.qubo_circuit(graph, steps, b, g)
        return float(np.real(bk.evaluate(circ.run().expectation(cuts))))

    e0 = expect(beta, gamma)
    eps = 1e-6
    dbeta, dgamma = np.zeros(steps), np.zeros(steps)
    for p in range(steps):
        d = np.zeros(steps)
        d[p] = eps
        dbeta[p] = (expect(beta + d, gamma) - expect(beta - d, gamma)) / (2 * eps)
        dgamma[p] = (expect(beta, gamma + d) - expect(beta, gamma - d)) / (2 * eps)
    ket = qf.qubo_circuit(graph, steps, beta, gamma).run()
    np.savez_compressed(os.path.join(HERE, 'qaoa.npz'), expectation=np.asarray([e0]), dbeta=dbeta, dgamma=dgamma,
                        ket=flat(ket.tensor), cuts=np.asarray(cuts).reshape(-1),
                        edges=np.asarray([[0, 1], [1, 2], [1, 3]]))
    # second instance: 6-node gnp graph of config C2
    g6 = nx.gnp_random_graph(6, 0.5, seed=0)
    np.random.seed(0)
    b6 = np.random.normal(0.5, 0.01, size=5)
    g6p = np.random.normal(0.5, 0.01, size=5)
    circ = qf.qubo_circuit(g6, 5, b6, g6p)
    ket6 = circ.run()
    e6 = float(np.real(bk.evaluate(ket6.expectation(qf.graph_cuts(g6)))))
    np.savez_compressed(os.path.join(HERE, 'qaoa6.npz'), expectation=np.asarray([e6]), beta=b6, gamma=g6p,
                        ket=flat(ket6.tensor), edges=np.asarray(list(g6.edges())),
                        cuts=np.asarray(qf.graph_cuts(g6)).reshape(-1))


if __name__ == '__main__':
    make_stdgates()
    make_tensormul()
    make_workloads()
    make_sampling()
    make_qaoa()
    print('golden fixtures written to', HERE)

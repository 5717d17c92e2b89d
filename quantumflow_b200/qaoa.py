"""QAOA workload builders (quantumflow/qaoa.py:22-81): the QUBO circuit and the diagonal cut-value cost."""
from typing import Sequence

import numpy as np

from .circuits import Circuit
from .stdgates import H, RX, RZ, ZZ

__all__ = ['qubo_circuit', 'graph_cuts']


def qubo_circuit(graph, steps: int, beta: Sequence, gamma: Sequence) -> Circuit:
    """H on every node; per step: ZZ(-w gamma_p / pi) per edge, RZ(node weight) per weighted node, RX(beta_p) per
    node. `graph` is a networkx graph; beta/gamma may hold torch tensors (autograd bridge)."""
    nodes = list(graph.nodes())
    circ = Circuit([H(q) for q in nodes])
    for p in range(steps):
        for a, b in graph.edges():
            weight = graph[a][b].get('weight', 1.0)
            circ += ZZ(-weight * gamma[p] / np.pi, a, b)
        for q in nodes:
            node_weight = graph.nodes[q].get('weight', None)
            if node_weight is not None:
                circ += RZ(node_weight, q)
        for q in nodes:
            circ += RX(beta[p], q)
    return circ


def graph_cuts(graph) -> np.ndarray:
    """Cut value of every bit assignment, as a float [2]*N array (axis i = node i). Vectorised over the 2^N
    assignments instead of the reference's per-entry Python loop; same values."""
    count = len(graph)
    index = np.arange(2 ** count, dtype=np.int64)
    cuts = np.zeros(2 ** count, dtype=np.double)
    for a, b in graph.edges():
        weight = graph[a][b].get('weight', 1)
        bit_a = (index >> (count - 1 - a)) & 1
        bit_b = (index >> (count - 1 - b)) & 1
        cuts += weight * (bit_a != bit_b)
    return cuts.reshape([2] * count)

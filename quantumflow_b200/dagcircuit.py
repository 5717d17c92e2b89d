"""Dependency-graph view of a circuit: depth, layers, components (quantumflow/dagcircuit.py:25-172).

Used here for the workload definitions ("depth-20" is `DAGCircuit.depth()`). The graph is kept as
per-qubit chains of elements in program order, which is all that depth / layers / neighbours need; program
order is a valid topological order, so iteration reproduces it.
"""
from typing import Dict, Iterable, Iterator, List, Optional

from .circuits import Circuit
from .ops import Channel, Gate, Operation
from .qubits import Qubit, Qubits
from .states import Density, State

__all__ = ['DAGCircuit']


class DAGCircuit(Operation):
    def __init__(self, elements: Iterable[Operation]) -> None:
        self._elements: List[Operation] = [e for e in elements if isinstance(e, Operation)]
        self._chains: Dict[Qubit, List[int]] = {}
        for n, elem in enumerate(self._elements):
            for q in elem.qubits:
                self._chains.setdefault(q, []).append(n)
        self._index = {id(e): n for n, e in enumerate(self._elements)}

    @property
    def qubits(self) -> Qubits:
        return tuple(sorted(self._chains))

    @property
    def qubit_nb(self) -> int:
        return len(self._chains)

    @property
    def H(self) -> 'DAGCircuit':
        return DAGCircuit(Circuit(self).H)

    def run(self, ket: State) -> State:
        return Circuit(self).run(ket)

    def evolve(self, rho: Density) -> Density:
        return Circuit(self).evolve(rho)

    def asgate(self) -> Gate:
        return Circuit(self).asgate()

    def aschannel(self) -> Channel:
        return Circuit(self).aschannel()

    def _levels(self, keep=None) -> List[int]:
        """Longest-chain level (0-based) of every kept element."""
        last: Dict[Qubit, int] = {}
        levels = [-1] * len(self._elements)
        for n, elem in enumerate(self._elements):
            if keep is not None and not keep(elem):
                continue
            lvl = 1 + max((last.get(q, -1) for q in elem.qubits), default=-1)
            levels[n] = lvl
            for q in elem.qubits:
                last[q] = lvl
        return levels

    def depth(self, local: bool = True) -> int:
        """Number of elements on the longest dependency chain; `local=False` ignores one-qubit elements."""
        keep = None if local else (lambda e: len(e.qubits) > 1)
        levels = self._levels(keep)
        return (max(levels) + 1) if levels else 0

    def size(self) -> int:
        return len(self._elements)

    def _component_labels(self) -> Dict[Qubit, Qubit]:
        parent = {q: q for q in self._chains}

        def find(q):
            while parent[q] != q:
                parent[q] = parent[parent[q]]
                q = parent[q]
            return q

        for elem in self._elements:
            qs = list(elem.qubits)
            for q in qs[1:]:
                parent[find(q)] = find(qs[0])
        return {q: find(q) for q in self._chains}

    def component_nb(self) -> int:
        return len(set(self._component_labels().values()))

    def components(self) -> List['DAGCircuit']:
        labels = self._component_labels()
        groups: Dict[Qubit, List[Operation]] = {}
        for elem in self._elements:
            groups.setdefault(labels[list(elem.qubits)[0]], []).append(elem)
        return [DAGCircuit(elems) for elems in groups.values()]

    def layers(self) -> Circuit:
        """Circuit of Circuits: layer d holds the elements whose longest chain from the inputs has length d."""
        levels = self._levels()
        layered: List[List[Operation]] = [[] for _ in range(self.depth())]
        for elem, lvl in zip(self._elements, levels):
            layered[lvl].append(elem)
        return Circuit([Circuit(layer) for layer in layered])

    def __iter__(self) -> Iterator[Operation]:
        return iter(self._elements)

    def _neighbour(self, elem: Operation, qubit: Optional[Qubit], step: int):
        n = self._index[id(elem)]
        for q in elem.qubits:
            if qubit is None or q == qubit:
                chain = self._chains[q]
                pos = chain.index(n) + step
                if pos < 0:
                    return ('in', q)
                if pos >= len(chain):
                    return ('out', q)
                return self._elements[chain[pos]]
        raise AssertionError('qubit not on element')

    def next_element(self, elem: Operation, qubit: Qubit = None):
        return self._neighbour(elem, qubit, +1)

    def prev_element(self, elem: Operation, qubit: Qubit = None):
        return self._neighbour(elem, qubit, -1)

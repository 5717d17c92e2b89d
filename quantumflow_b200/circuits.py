"""Circuits: ordered lists of Operations, plus the standard library circuits.

Behavioural contract: quantumflow/circuits.py:34-372. `Circuit.run` / `Circuit.evolve` return exactly what the
reference's element-by-element fold returns, but runs of gates are handed to the planner
(quantumflow_b200/planner.py) and executed as a few tiled sweeps over an engine-owned copy of the state: the
caller's State is never mutated (gates and states stay immutable by convention), the copy is updated in place.
Elements that consult the RNG or classical memory (Measure, Kraus, Reset, If ...) are barriers and run through
their own `.run` / `.evolve`.
"""
import os
from collections import defaultdict
from itertools import chain
from math import pi
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple, Type

import torch

from . import backend as bk
from .gates import control_gate, identity_gate
from .ops import Channel, Gate, Operation
from .qubits import Qubit, Qubits
from .states import Density, State, zero_state
from .stdgates import CCNOT, CNOT, CPHASE, H, SWAP, T, TY, TZ, X

__all__ = ['Circuit', 'count_operations', 'map_gate', 'qft_circuit', 'reversal_circuit', 'control_circuit',
           'ccnot_circuit', 'zyz_circuit', 'phase_estimation_circuit', 'addition_circuit', 'ghz_circuit']

# states smaller than this are latency-bound; they go through the one-gate kernels
PLANNER_MIN_BITS = 6
# a run of gates shorter than this is not worth a plan
PLANNER_MIN_OPS = 2


def _plannable_gate(elem: Operation) -> bool:
    if not isinstance(elem, Gate):
        return False
    return not bool(getattr(elem.tensor, 'requires_grad', False))


class Circuit(Operation):
    """A sequence of Operations (gates, channels, measurements, nested circuits)."""

    def __init__(self, elements: Iterable[Operation] = None) -> None:
        self.elements = list(elements) if elements is not None else []
        self._plan_cache: Dict = {}

    # -- container protocol ---------------------------------------------------------------------------
    def add(self, other: 'Circuit') -> 'Circuit':
        return Circuit(self.elements + other.elements)

    def extend(self, other: Operation) -> None:
        if isinstance(other, Circuit):
            self.elements.extend(other.elements)
        else:
            self.elements.append(other)

    def __add__(self, other: 'Circuit') -> 'Circuit':
        return self.add(other)

    def __iadd__(self, other: Operation) -> 'Circuit':
        self.extend(other)
        return self

    def __iter__(self) -> Iterator[Operation]:
        return iter(self.elements)

    def size(self) -> int:
        return len(self.elements)

    @property
    def qubits(self) -> Qubits:
        """Sorted union of the element qubits."""
        return tuple(sorted({q for elem in self.elements for q in elem.qubits}))

    # -- execution ------------------------------------------------------------------------------------
    def _flat_elements(self) -> List[Operation]:
        flat: List[Operation] = []
        for elem in self.elements:
            if isinstance(elem, Circuit):
                flat.extend(elem._flat_elements())
            else:
                flat.append(elem)
        return flat

    def _segments(self, flat: List[Operation], key: Tuple, nbits: int, make_bitops):
        """Cached planner output for one run of gates (identified by the element identities)."""
        from . import planner
        cache = self.__dict__.setdefault('_plan_cache', {})
        fingerprint = (key, nbits, tuple(id(e) for e in flat))
        hit = cache.get(fingerprint)
        if hit is None:
            if len(cache) > 8:
                cache.clear()
            hit = (planner.build_segments(nbits, make_bitops(flat)), list(flat))   # keep elements alive
            cache[fingerprint] = hit
        return hit[0]

    @staticmethod
    def _execute(segments, tensor) -> None:
        """Run planner segments in place on an engine-owned tensor."""
        from . import engine
        for seg in segments:
            if seg.kind == 'plan':
                if seg.uploaded is None:
                    seg.uploaded = engine.UploadedPlan(seg.blob)
                seg.uploaded.launch(tensor)
            else:
                engine.apply_operator(tensor, seg.mat, seg.bits, inplace=True)

    def run(self, ket: State = None, _owned: bool = False) -> State:
        """Apply the circuit to a state (default |0...0> on the circuit's qubits). `_owned` (engine-internal
        callers such as Program.run): the buffer of `ket` belongs to the caller's own intermediate state and may
        be updated in place."""
        owned = ket is None or _owned          # an engine-created buffer may be updated in place
        if ket is None:
            ket = zero_state(qubits=self.qubits)
        flat = self._flat_elements()
        count = ket.qubit_nb
        i = 0
        while i < len(flat):
            # differentiable runs on small states: the whole run is one launch and one autograd node each way
            # (autograd.run_small_circuit; the QAOA gradient step of examples/qaoa_maxcut.py)
            g_end = i
            while g_end < len(flat) and isinstance(flat[g_end], Gate):
                g_end += 1
            if g_end - i >= 2 and torch.is_grad_enabled() and os.environ.get('QFB_SMALL_CIRCUIT', '1') != '0' and (
                    getattr(ket.tensor, 'requires_grad', False)
                    or any(getattr(e.tensor, 'requires_grad', False) for e in flat[i:g_end])):
                from . import autograd
                if autograd.small_circuit_shape_ok(flat[i:g_end], count):
                    where = {q: count - 1 - w for w, q in enumerate(ket.qubits)}
                    tensor = autograd.run_small_circuit(ket.tensor, flat[i:g_end],
                                                        lambda g: [where[q] for q in g.qubits])
                    if tensor is not None:           # None: a gate is not unitary -> gate by gate below
                        ket = State(tensor, ket.qubits, ket.memory)
                        owned = True
                        i = g_end
                        continue
            j = i
            while j < len(flat) and _plannable_gate(flat[j]):
                j += 1
            if (j - i >= PLANNER_MIN_OPS and count >= PLANNER_MIN_BITS
                    and not getattr(ket.tensor, 'requires_grad', False)):
                qubits = ket.qubits

                def bitops(gates, qubits=qubits, count=count):
                    return [(g.matrix(), [count - 1 - qubits.index(q) for q in g.qubits]) for g in gates]

                segments = self._segments(flat[i:j], ('run', tuple(qubits)), count, bitops)
                tensor = ket.tensor if owned else ket.tensor.clone()
                self._execute(segments, tensor)
                ket = State(tensor, ket.qubits, ket.memory)
                owned = True
                i = j
            else:
                for elem in flat[i:max(j, i + 1)]:
                    before = ket.tensor.data_ptr()
                    ket = elem.run(ket)
                    owned = owned or ket.tensor.data_ptr() != before
                i = max(j, i + 1)
        return ket

    def run_pipelined(self, host_inputs, host_outputs, qubits: Qubits = None, depth: int = 3) -> None:
        """Apply the circuit to a stream of states that live in HOST memory (engine API, no reference equivalent).

        host_inputs / host_outputs: equally long sequences of pinned, contiguous complex128 host tensors with
        2^n elements each (n = number of qubits; an input may appear several times). State i is copied to the
        device, run through the circuit's plan and copied back to host_outputs[i]. Three streams and `depth`
        device buffers form a pipeline: the upload of state i+1 and the download of state i-1 run under the
        sweeps of state i, so the sustained cost per state is max(upload, sweeps, download) instead of their sum
        (PCIe is full duplex). Only gate-only circuits (what the planner accepts) are supported."""
        import torch
        qubits = tuple(self.qubits if qubits is None else qubits)
        count = len(qubits)
        flat = self._flat_elements()
        if count < PLANNER_MIN_BITS or not all(_plannable_gate(e) for e in flat):
            raise ValueError('run_pipelined needs a gate-only circuit on at least {} qubits'.format(PLANNER_MIN_BITS))
        if len(host_inputs) != len(host_outputs):
            raise ValueError('host_inputs and host_outputs differ in length')
        for t in list(host_inputs) + list(host_outputs):
            if t.is_cuda or not t.is_pinned() or t.dtype != torch.complex128 or t.numel() != 1 << count \
                    or not t.is_contiguous():
                raise ValueError('run_pipelined needs pinned contiguous complex128 host tensors of 2^n elements')

        def bitops(gates):
            return [(g.matrix(), [count - 1 - qubits.index(q) for q in g.qubits]) for g in gates]

        segments = self._segments(flat, ('run', qubits), count, bitops)
        dev = torch.device('cuda', torch.cuda.current_device())
        depth = max(1, min(int(depth), len(host_inputs)))
        buffers = [torch.empty(1 << count, dtype=torch.complex128, device=dev) for _ in range(depth)]
        up, work, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(dev))
        for st in (up, work, down):
            st.wait_event(start)
        downloaded = [None] * depth             # event: the buffer's previous result has left the device
        last = None
        for i, (src, dst) in enumerate(zip(host_inputs, host_outputs)):
            buf = buffers[i % depth]
            with torch.cuda.stream(up):
                if downloaded[i % depth] is not None:
                    up.wait_event(downloaded[i % depth])
                buf.copy_(src.reshape(-1), non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(up)
            with torch.cuda.stream(work):
                work.wait_event(uploaded)
                self._execute(segments, buf)
                computed = torch.cuda.Event()
                computed.record(work)
            with torch.cuda.stream(down):
                down.wait_event(computed)
                dst.reshape(-1).copy_(buf, non_blocking=True)
                last = torch.cuda.Event()
                last.record(down)
                downloaded[i % depth] = last
        if last is not None:
            torch.cuda.current_stream(dev).wait_event(last)
            last.synchronize()                   # the host buffers are valid when the call returns

    def evolve(self, rho: Density = None, _owned: bool = False) -> Density:
        """Apply the circuit to a density matrix (default |0...0><0...0|); `_owned` as in run()."""
        owned = rho is None or _owned
        if rho is None:
            rho = zero_state(qubits=self.qubits).asdensity()
        flat = self._flat_elements()
        count = rho.qubit_nb
        i = 0

        def plannable(elem):
            if _plannable_gate(elem):
                return True
            if isinstance(elem, Channel):
                return elem.qubit_nb == 1 and not getattr(elem.tensor, 'requires_grad', False)
            superop = getattr(elem, 'superoperator_matrix', None)
            return superop is not None and elem.qubit_nb == 1

        while i < len(flat):
            j = i
            while j < len(flat) and plannable(flat[j]):
                j += 1
            if j - i >= PLANNER_MIN_OPS and 2 * count >= PLANNER_MIN_BITS:
                qubits = rho.qubits

                def bitops(elems, qubits=qubits, count=count):
                    out = []
                    for e in elems:
                        where = [qubits.index(q) for q in e.qubits]
                        ket_bits = [2 * count - 1 - w for w in where]
                        bra_bits = [count - 1 - w for w in where]
                        if isinstance(e, Gate):
                            mat = e.matrix()
                            out.append((mat, ket_bits))
                            out.append((mat.conj(), bra_bits))
                        elif isinstance(e, Channel):
                            dim = 4 ** e.qubit_nb
                            out.append((bk.evaluate(e.tensor).reshape(dim, dim), ket_bits + bra_bits))
                        else:
                            out.append((e.superoperator_matrix(), ket_bits + bra_bits))
                    return out

                segments = self._segments(flat[i:j], ('evolve', tuple(qubits)), 2 * count, bitops)
                tensor = rho.tensor if owned else rho.tensor.clone()
                self._execute(segments, tensor)
                # Kraus.evolve drops classical memory in the reference (channels.py:85)
                memory = rho.memory
                if any(hasattr(e, 'superoperator_matrix') for e in flat[i:j]):
                    memory = None
                rho = Density(tensor, rho.qubits, memory)
                owned = True
                i = j
            else:
                for elem in flat[i:max(j, i + 1)]:
                    before = rho.tensor.data_ptr()
                    rho = elem.evolve(rho)
                    owned = owned or rho.tensor.data_ptr() != before
                i = max(j, i + 1)
        return rho

    def asgate(self) -> Gate:
        gate = identity_gate(self.qubits)
        for elem in self.elements:
            gate = elem.asgate() @ gate
        return gate

    def aschannel(self) -> Channel:
        chan = identity_gate(self.qubits).aschannel()
        for elem in self.elements:
            chan = elem.aschannel() @ chan
        return chan

    @property
    def H(self) -> 'Circuit':
        """Reversed circuit of element conjugates (the inverse when every element is unitary)."""
        return Circuit([elem.H for elem in reversed(self.elements)])

    def __str__(self) -> str:
        return '\n'.join(str(elem) for elem in self.elements)


def count_operations(elements: Iterable[Operation]) -> Dict[Type[Operation], int]:
    tally: Dict[Type[Operation], int] = defaultdict(int)
    for elem in elements:
        tally[type(elem)] += 1
    return dict(tally)


def map_gate(gate: Gate, args: Sequence[Qubits]) -> Circuit:
    """One relabelled copy of `gate` per qubit tuple in `args`."""
    circ = Circuit()
    for qubits in args:
        circ += gate.relabel(qubits)
    return circ


def qft_circuit(qubits: Qubits) -> Circuit:
    """Quantum Fourier transform: H then controlled phases pi/2^d per qubit, then bit reversal."""
    count = len(qubits)
    circ = Circuit()
    for a in range(count):
        circ += H(qubits[a])
        for b in range(a + 1, count):
            circ += CPHASE(pi / 2 ** (b - a), qubits[b], qubits[a])
    circ.extend(reversal_circuit(qubits))
    return circ


def reversal_circuit(qubits: Qubits) -> Circuit:
    count = len(qubits)
    return Circuit([SWAP(qubits[a], qubits[count - 1 - a]) for a in range(count // 2)])


def control_circuit(controls: Qubits, gate: Gate) -> Circuit:
    """`gate` controlled on all of `controls` (Barenco et al. 1995, sec 7.2; quadratic gate count)."""
    circ = Circuit()
    if len(controls) == 1:
        c = controls[0]
        if isinstance(gate, X):
            circ += CNOT(c, gate.qubits[0])
        else:
            circ += control_gate(c, gate)
        return circ
    last, rest = controls[-1:], controls[:-1]
    circ += control_circuit(last, gate ** 0.5)
    circ += control_circuit(rest, X(controls[-1]))
    circ += control_circuit(last, gate ** -0.5)
    circ += control_circuit(rest, X(controls[-1]))
    circ += control_circuit(rest, gate ** 0.5)
    return circ


def ccnot_circuit(qubits: Qubits) -> Circuit:
    """Toffoli from 6 CNOTs, Hadamards and T gates (Nielsen & Chuang)."""
    if len(qubits) != 3:
        raise ValueError('Expected 3 qubits')
    a, b, c = qubits
    return Circuit([H(c), CNOT(b, c), T(c).H, CNOT(a, c), T(c), CNOT(b, c), T(c).H, CNOT(a, c), T(b), T(c), H(c),
                    CNOT(a, b), T(a), T(b).H, CNOT(a, b)])


def zyz_circuit(t0: float, t1: float, t2: float, q0: Qubit) -> Circuit:
    return Circuit([TZ(t0, q0), TY(t1, q0), TZ(t2, q0)])


def phase_estimation_circuit(gate: Gate, outputs: Qubits) -> Circuit:
    """Phase estimation of an eigenphase of `gate` into the `outputs` register (inverse QFT at the end)."""
    circ = Circuit()
    circ += map_gate(H(), list(zip(outputs)))
    for cq in reversed(outputs):
        circ += control_gate(cq, gate)
        gate = gate @ gate
    circ += qft_circuit(outputs).H
    return circ


def addition_circuit(addend0: Qubits, addend1: Qubits, carry: Qubits) -> Circuit:
    """Cuccaro ripple-carry adder; the sum replaces addend1, carry = (carry in, carry out)."""
    if len(addend0) != len(addend1):
        raise ValueError('Number of addend qubits must be equal')
    if len(carry) != 2:
        raise ValueError('Expected 2 carry qubits')

    def majority(a, b, c):
        return Circuit([CNOT(c, b), CNOT(c, a), CCNOT(a, b, c)])

    def unmajority_add(a, b, c):
        return Circuit([CCNOT(a, b, c), CNOT(c, a), CNOT(a, b)])

    wires = [carry[0]] + list(chain.from_iterable(zip(reversed(addend1), reversed(addend0)))) + [carry[1]]
    starts = range(0, len(wires) - 3, 2)
    circ = Circuit()
    for n in starts:
        circ += majority(*wires[n:n + 3])
    circ += CNOT(wires[-2], wires[-1])
    for n in reversed(starts):
        circ += unmajority_add(*wires[n:n + 3])
    return circ


def ghz_circuit(qubits: Qubits) -> Circuit:
    """H on the first qubit then a CNOT chain."""
    circ = Circuit([H(qubits[0])])
    for a in range(len(qubits) - 1):
        circ += CNOT(qubits[a], qubits[a + 1])
    return circ

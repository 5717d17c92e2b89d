"""Batched stochastic trajectories: B = 2^b pure states of N qubits in ONE device buffer.

Reference: `Kraus.run` / `UnitaryMixture.run` (quantumflow/channels.py:70-77, 119-125) unravel a channel on ONE state
per call -- every branch K_k psi is computed, the branch norms become probabilities, `np.random.choice` picks one,
the state is renormalised. Monte-Carlo noise needs many such trajectories; here they share the buffer
(trajectory index = the top b index bits) and every operation acts on all of them at once:

  gates                 the batch is an (N+b)-bit state whose top bits no gate touches: the planner and the sweep
                        kernels run all trajectories in the same sweeps (csrc/qfb_sweep.cu, qfb_jit.cu)
  1-qubit Kraus channel two passes over the batch whatever the number of operators: qfb_batch_rho1 (per-trajectory
                        reduced density of the qubit -> every branch probability w_k tr(K_k rho K_k^dagger) on the
                        host) and qfb_batch_apply1 (the drawn branch, divided by its norm, as a 2x2 operator PER
                        TRAJECTORY selected by the trajectory bits of the index)
  mixture of unitaries  (Depolarizing, Dephasing): probabilities are the weights, so only the second pass
  anything else         (multi-qubit Kraus, Measure, ...) falls back to the per-state implementation, trajectory
                        by trajectory, still on the device

RNG contract: the draws are the ones B sequential calls make, in trajectory order, from numpy's global stream --
`np.random.choice(n, p=p)` consumes exactly one `random_sample()` and returns searchsorted(cumsum(p) / cumsum(p)[-1],
u, side='right'); `run` draws the B uniforms of one channel with one call. A circuit is
therefore equivalent to the reference loop `for op in circuit: for t in range(B): ket[t] = op.run(ket[t])`
(tests/golden/make_golden_trajectories.py produces exactly that with the reference; tests/test_gpu_trajectories.py).
Replicas over GPUs: trajectories never interact, so each rank of a multi-GPU job holds its own StateBatch (seeded
per rank); no collective is involved.
"""
from typing import List, Sequence

import numpy as np
import torch

from . import _lib, engine, planner
from .channels import Kraus, UnitaryMixture
from .ops import Gate
from .qubits import Qubits
from .states import State

__all__ = ['StateBatch', 'choice_indices']


def choice_indices(probs: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """What `[np.random.choice(n, p=probs[t]) for t in range(B)]` returns when its B uniform draws are `uniforms`:
    row-wise searchsorted(cdf / cdf[-1], u, side='right')."""
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64), axis=1)
    cdf /= cdf[:, -1:]
    return (cdf <= np.asarray(uniforms, dtype=np.float64)[:, None]).sum(axis=1).astype(np.int64)


class StateBatch:
    """2^batch_bits pure states on the same qubits, stored back to back in one complex128 device buffer."""

    def __init__(self, tensor: torch.Tensor, qubits: Qubits, batch_bits: int) -> None:
        self.qubits = tuple(qubits)
        self.qubit_nb = len(self.qubits)
        self.batch_bits = int(batch_bits)
        self.tensor = tensor.reshape(-1)
        if not (self.tensor.is_cuda and self.tensor.dtype == torch.complex128 and self.tensor.is_contiguous()):
            raise ValueError('a StateBatch lives in HBM as a contiguous complex128 tensor')
        if self.tensor.numel() != 1 << (self.qubit_nb + self.batch_bits):
            raise ValueError('Incompatibility between tensor and qubits')
        self._plans = {}

    @property
    def size(self) -> int:
        return 1 << self.batch_bits

    @classmethod
    def zeros(cls, size: int, qubits: Qubits) -> 'StateBatch':
        """`size` (a power of two) copies of |0...0>."""
        qubits = tuple(range(qubits)) if isinstance(qubits, int) else tuple(qubits)
        b = int(size).bit_length() - 1
        if (1 << b) != size:
            raise ValueError('the number of trajectories must be a power of two')
        from .backend import device
        t = torch.zeros(size, 1 << len(qubits), dtype=torch.complex128, device=device())
        t[:, 0] = 1.0
        return cls(t, qubits, b)

    @classmethod
    def from_states(cls, states: Sequence[State]) -> 'StateBatch':
        b = len(states).bit_length() - 1
        if (1 << b) != len(states) or any(s.qubits != states[0].qubits for s in states):
            raise ValueError('need a power-of-two number of states on the same qubits')
        return cls(torch.stack([s.tensor.reshape(-1) for s in states]).contiguous(), states[0].qubits, b)

    def state(self, t: int) -> State:
        n = 1 << self.qubit_nb
        return State(self.tensor[t * n:(t + 1) * n].clone().reshape([2] * self.qubit_nb), self.qubits)

    def asarray(self) -> np.ndarray:
        return self.tensor.cpu().numpy().reshape(self.size, 1 << self.qubit_nb)

    # ---- operations ------------------------------------------------------------------------------------
    def _bit(self, qubit) -> int:
        return self.qubit_nb - 1 - self.qubits.index(qubit)

    def _reduced_density(self, bit: int) -> np.ndarray:
        """[B, 4] = (p0, p1, Re rho01, Im rho01) of the qubit at index bit `bit`, per trajectory."""
        lib = _lib.load()
        out = torch.empty(self.size * 4, dtype=torch.float64, device=self.tensor.device)
        nbytes = lib.qfb_batch_rho1_workspace(self.qubit_nb, self.batch_bits)
        work = torch.empty(max(1, nbytes // 8), dtype=torch.float64, device=self.tensor.device)
        _lib.check(lib.qfb_batch_rho1(self.tensor.data_ptr(), self.qubit_nb, self.batch_bits, bit, out.data_ptr(),
                                      work.data_ptr(), nbytes, engine._stream()))
        return out.cpu().numpy().reshape(self.size, 4)

    def _apply_per_trajectory(self, bit: int, mats: np.ndarray) -> None:
        lib = _lib.load()
        table = torch.from_numpy(np.ascontiguousarray(mats, dtype=np.complex128).reshape(self.size, 4)).to(
            self.tensor.device)
        _lib.check(lib.qfb_batch_apply1(self.tensor.data_ptr(), self.qubit_nb, self.batch_bits, bit,
                                        table.data_ptr(), engine._stream()))
        torch.cuda.current_stream().synchronize()      # `table` must outlive the kernel

    def _run_gates(self, gates: List[Gate]) -> None:
        key = tuple(id(g) for g in gates)
        segments = self._plans.get(key)
        if segments is None:
            nbits = self.qubit_nb + self.batch_bits
            bitops = [(g.matrix(), [self._bit(q) for q in g.qubits]) for g in gates]
            if nbits >= planner.MIN_TILE_BITS and len(gates) >= 2:
                segments = planner.build_segments(nbits, bitops)
            else:
                segments = [planner.Segment('op', mat=m, bits=b, nsweeps=1, nops=1) for m, b in bitops]
            self._plans[key] = (segments, list(gates))
        else:
            segments = segments[0]
        from .circuits import Circuit
        Circuit._execute(segments, self.tensor)

    def _run_kraus(self, kraus: Kraus) -> None:
        ops = list(kraus.operators)
        weights = np.asarray(kraus.weights, dtype=np.float64)
        single = all(op.qubit_nb == 1 and op.qubits == ops[0].qubits for op in ops)
        if not single:
            # per-state implementation, trajectory by trajectory (same draws in the same order)
            n = 1 << self.qubit_nb
            for t in range(self.size):
                self.tensor[t * n:(t + 1) * n] = kraus.run(self.state(t)).tensor.reshape(-1)
            return
        bit = self._bit(ops[0].qubits[0])
        mats = np.stack([np.asarray(op.matrix(), dtype=np.complex128).reshape(2, 2) for op in ops])
        if isinstance(kraus, UnitaryMixture):
            # UnitaryMixture.asgate: np.random.choice(operators, p=weights), no renormalisation
            probs = np.broadcast_to(weights, (self.size, len(ops)))
            pick = choice_indices(probs, np.random.random_sample(self.size))
            chosen = mats[pick]
        else:
            rho = self._reduced_density(bit)
            gram = np.einsum('kji,kjl->kil', mats.conj(), mats)              # K_k^dagger K_k
            r01 = rho[:, 2] + 1j * rho[:, 3]
            branch = (gram[None, :, 0, 0].real * rho[:, None, 0] + gram[None, :, 1, 1].real * rho[:, None, 1] +
                      2.0 * (gram[None, :, 1, 0] * r01[:, None]).real)       # |K_k psi_t|^2
            probs = branch * weights[None, :]
            pick = choice_indices(probs, np.random.random_sample(self.size))
            norm = branch[np.arange(self.size), pick]
            chosen = mats[pick] / np.sqrt(norm)[:, None, None]
        self._apply_per_trajectory(bit, chosen)

    def run(self, operation) -> 'StateBatch':
        """Apply a gate, a Kraus operation or a circuit of them to every trajectory, IN PLACE; returns self."""
        elements = list(getattr(operation, 'elements', [operation]))
        flat: List[object] = []
        for elem in elements:
            flat.extend(elem._flat_elements() if hasattr(elem, '_flat_elements') else [elem])
        pending: List[Gate] = []
        for elem in flat:
            if isinstance(elem, Gate):
                pending.append(elem)
                continue
            if pending:
                self._run_gates(pending)
                pending = []
            if isinstance(elem, Kraus):
                self._run_kraus(elem)
            else:
                n = 1 << self.qubit_nb
                for t in range(self.size):
                    self.tensor[t * n:(t + 1) * n] = elem.run(self.state(t)).tensor.reshape(-1)
        if pending:
            self._run_gates(pending)
        return self

    def norms(self) -> np.ndarray:
        rho = self._reduced_density(0)
        return rho[:, 0] + rho[:, 1]

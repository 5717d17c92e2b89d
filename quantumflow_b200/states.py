"""Pure and mixed states resident in HBM.

Behavioural contract: quantumflow/states.py:29-357. A `State` owns a complex128 CUDA tensor of shape [2]*N
(16 * 2^N bytes); a `Density` one of shape [2]*2N, ket axes first. All arithmetic on them is a libqfb200 kernel
(norm, probabilities, expectation, outer product ...). Sampling keeps numpy's global RandomState as the shared
RNG stream: probabilities come from the device kernel, the draw itself is the very numpy call the reference
makes (states.py:121-129, 149-160), so the stream advances identically.
"""
from collections import ChainMap, defaultdict
from functools import reduce
from math import sqrt
from typing import Any, Dict, Sequence, TextIO, Union

import numpy as np

from . import backend as bk
from . import engine
from .cbits import Addr
from .qubits import Qubits, QubitVector, outer_product, qubits_count_tuple

__all__ = ['State', 'ghz_state', 'join_states', 'print_probabilities', 'print_state', 'random_state', 'w_state',
           'zero_state', 'Density', 'mixed_density', 'random_density', 'join_densities']


def _basis_bits(index: int, count: int) -> np.ndarray:
    """Bits of a flat basis index, qubit 0 first (axis 0 is the most significant bit)."""
    return np.array([(index >> (count - 1 - i)) & 1 for i in range(count)], dtype=np.int64)


class State:
    """State vector of N qubits, stored in HBM."""

    _RANK = 1

    def __init__(self, tensor: bk.TensorLike, qubits: Qubits = None, memory: Dict[Addr, Any] = None) -> None:
        if qubits is None:
            tensor = bk.astensorproduct(tensor)
            qubits = range(bk.rank(tensor) // self._RANK)
        self.vec = QubitVector(tensor, qubits, resident=True)
        self._memory = memory if memory is not None else {}

    # -- structure ------------------------------------------------------------------------------------
    @property
    def tensor(self) -> bk.BKTensor:
        return self.vec.tensor

    @property
    def qubits(self) -> Qubits:
        return self.vec.qubits

    @property
    def qubit_nb(self) -> int:
        return self.vec.qubit_nb

    @property
    def memory(self) -> dict:
        """Classical memory; a fresh defaultdict(int) on every access (states.py:78-79)."""
        return defaultdict(int, self._memory)

    def update(self, memory: Dict[Addr, Any]) -> 'State':
        merged = self.memory
        merged.update(memory)
        return type(self)(self.tensor, self.qubits, merged)

    @property
    def cbits(self) -> Sequence[Addr]:
        return tuple(sorted(addr for addr in self._memory if addr.dtype == 'BIT'))

    @property
    def cbit_nb(self) -> int:
        return len(self.cbits)

    def relabel(self, qubits: Qubits) -> 'State':
        return type(self)(self.vec.tensor, qubits, self._memory)

    def permute(self, qubits: Qubits) -> 'State':
        vec = self.vec.permute(qubits)
        return type(self)(vec.tensor, vec.qubits, self._memory)

    # -- read-out -------------------------------------------------------------------------------------
    def norm(self) -> bk.BKTensor:
        """<psi|psi> (squared norm; reference convention)."""
        return self.vec.norm()

    def normalize(self) -> 'State':
        if self.tensor.requires_grad:
            tensor = self.tensor / bk.ccast(bk.sqrt(self.norm()))
        else:
            tensor = engine.normalize_by_norm2(self.tensor)
        return State(tensor, self.qubits, self._memory)

    def probabilities(self) -> bk.BKTensor:
        """|amplitude|^2 as a float64 [2]*N device tensor."""
        if self.tensor.requires_grad:
            from .autograd import probabilities_autograd
            return probabilities_autograd(self.tensor)
        return engine.probabilities(self.tensor)

    def _host_probabilities(self) -> np.ndarray:
        return np.real(bk.evaluate(self.probabilities()))

    def sample(self, trials: int) -> np.ndarray:
        """Counts of each basis outcome over `trials` measurements (numpy global RNG, one multinomial call)."""
        probs = self._host_probabilities()
        counts = np.random.multinomial(trials, probs.ravel())
        return counts.reshape(probs.shape)

    def expectation(self, diag_hermitian: bk.TensorLike, trials: int = None) -> bk.BKTensor:
        """Expectation of a Hermitian operator that is diagonal in the computational basis."""
        if trials is not None:
            freq = self.sample(trials) / trials
            diag = np.real(np.asarray(bk.evaluate(bk.astensorproduct(diag_hermitian))))
            return bk.fcast(np.sum(diag * freq))
        if self.tensor.requires_grad:
            from .autograd import expectation_autograd
            return expectation_autograd(self.tensor, diag_hermitian)
        import torch
        if isinstance(diag_hermitian, torch.Tensor):
            diag = diag_hermitian
            if diag.is_complex():
                diag = diag.real
            diag = diag.to(device=self.tensor.device, dtype=bk.FTYPE)
        else:
            diag = torch.from_numpy(np.ascontiguousarray(np.real(np.asarray(diag_hermitian)),
                                                         dtype=np.float64)).to(self.tensor.device)
        if self._RANK == 1:
            return engine.expectation_diag(self.tensor, diag.reshape(-1))
        probs = bk.real(self.probabilities())
        return bk.sum(diag.reshape(probs.shape) * probs)

    def measure(self) -> np.ndarray:
        """One computational-basis outcome as an array of N bits (numpy global RNG, one `choice` call)."""
        probs = self._host_probabilities()
        outcome = np.random.choice(probs.size, p=probs.ravel())
        return _basis_bits(int(outcome), self.qubit_nb)

    def asdensity(self) -> 'Density':
        """|psi><psi| built on the device (one outer-product sweep)."""
        matrix = engine.outer(self.tensor, self.tensor, conj_second=True)
        return Density(matrix, self.qubits, self._memory)

    def __str__(self) -> str:
        terms = []
        for index, amplitude in np.ndenumerate(self.vec.asarray()):
            if np.isclose(amplitude, 0.0):
                continue
            if len(terms) > 64:
                terms.append('...')
                break
            ket = ''.join(str(b) for b in index)
            terms.append('({c.real:0.04g}{c.imag:+0.04g}i) |{k}>'.format(c=amplitude, k=ket))
        return ' + '.join(terms)


def _basis_state(qubits: Union[int, Qubits], entries: Dict[int, complex]) -> State:
    """State with the given flat-index amplitudes, built directly in HBM."""
    import torch
    count, qubits = qubits_count_tuple(qubits)
    tensor = torch.zeros(1 << count, dtype=bk.CTYPE, device=bk.device())
    if entries:
        idx = torch.tensor(list(entries.keys()), dtype=torch.long, device=tensor.device)
        val = torch.tensor(list(entries.values()), dtype=bk.CTYPE, device=tensor.device)
        tensor[idx] = val
    return State(tensor.reshape([2] * count), qubits)


def zero_state(qubits: Union[int, Qubits]) -> State:
    """|0...0>"""
    return _basis_state(qubits, {0: 1.0})


def w_state(qubits: Union[int, Qubits]) -> State:
    """Equal superposition of all single-excitation basis states."""
    count, qubits = qubits_count_tuple(qubits)
    return _basis_state(qubits, {1 << (count - 1 - i): 1 / sqrt(count) for i in range(count)})


def ghz_state(qubits: Union[int, Qubits]) -> State:
    """(|0...0> + |1...1>)/sqrt(2)"""
    count, qubits = qubits_count_tuple(qubits)
    return _basis_state(qubits, {0: 1 / sqrt(2), (1 << count) - 1: 1 / sqrt(2)})


def random_state(qubits: Union[int, Qubits]) -> State:
    """Gaussian random state; draws real parts then imaginary parts from numpy's global RNG (states.py:214-219)."""
    count, qubits = qubits_count_tuple(qubits)
    re = np.random.normal(size=([2] * count))
    im = np.random.normal(size=([2] * count))
    return State(re + 1j * im, qubits).normalize()


def join_states(*states: State) -> State:
    vec = reduce(outer_product, [ket.vec for ket in states])
    return State(vec.tensor, vec.qubits)


def print_state(state: State, file: TextIO = None) -> None:
    for index, amplitude in np.ndenumerate(state.vec.asarray()):
        print(''.join(str(b) for b in index), ':', amplitude, file=file)


def print_probabilities(state: State, ndigits: int = 4, file: TextIO = None) -> None:
    probs = np.real(bk.evaluate(state.probabilities()))
    for index, prob in np.ndenumerate(probs):
        prob = round(float(prob), ndigits)
        if prob == 0.0:
            continue
        print(''.join(str(b) for b in index), ':', prob, file=file)


class Density(State):
    """Density matrix of N qubits: [2]*2N tensor in HBM, ket (row) axes first."""

    _RANK = 2

    def trace(self) -> bk.BKTensor:
        return self.vec.trace()

    def partial_trace(self, qubits: Qubits) -> 'Density':
        vec = self.vec.partial_trace(qubits)
        return Density(vec.tensor, vec.qubits, self._memory)

    def normalize(self) -> 'Density':
        """rho / tr(rho), trace kept on the device."""
        tensor = engine.divide_by_device_scalar(self.tensor, self.trace())
        return Density(tensor, self.qubits, self._memory)

    def probabilities(self) -> bk.BKTensor:
        """Diagonal of rho as a (complex-typed, like the reference) [2]*N tensor."""
        return bk.productdiag(self.tensor)

    def asoperator(self) -> bk.BKTensor:
        return self.vec.flatten()

    def asdensity(self) -> 'Density':
        return self


def mixed_density(qubits: Union[int, Qubits]) -> Density:
    """The completely mixed state I / 2^N."""
    count, qubits = qubits_count_tuple(qubits)
    return Density(np.eye(2 ** count) / 2 ** count, qubits)


def random_density(qubits: Union[int, Qubits]) -> Density:
    """Hilbert-Schmidt ensemble: G G^dagger / tr(.) with G Ginibre (states.py:329-345; same RNG call order)."""
    count, qubits = qubits_count_tuple(qubits)
    shape = (2 ** count, 2 ** count)
    ginibre = (np.random.normal(size=shape) + 1j * np.random.normal(size=shape)) / np.sqrt(2.0)
    matrix = ginibre @ ginibre.conj().T
    matrix /= np.trace(matrix)
    return Density(matrix, qubits=qubits)


def join_densities(*densities: Density) -> Density:
    vec = reduce(outer_product, [rho.vec for rho in densities])
    memory = dict(ChainMap(*[rho.memory for rho in densities]))
    return Density(vec.tensor, vec.qubits, memory)

"""ctypes wrapper of oracle/libqforacle.so (C restatement of tensormul; TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'libqforacle.so')
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'qf_oracle_c.c')
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, '-B', 'libqforacle.so'], check=True, capture_output=True)
    return _PATH


def load():
    global _LIB
    if _LIB is None:
        build()
        lib = ctypes.CDLL(_PATH)
        lib.qfo_apply_dense.restype = ctypes.c_int
        lib.qfo_apply_dense.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int)]
        lib.qfo_norm2.restype = ctypes.c_double
        lib.qfo_norm2.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.qfo_max_threads.restype = ctypes.c_int
        lib.qfo_set_threads.argtypes = [ctypes.c_int]
        _LIB = lib
    return _LIB


def apply_dense(state: np.ndarray, mat: np.ndarray, bits) -> None:
    """In-place on a C-contiguous complex128 vector of 2^n amplitudes."""
    lib = load()
    assert state.dtype == np.complex128 and state.flags['C_CONTIGUOUS']
    nbits = int(np.log2(state.size))
    mat = np.ascontiguousarray(mat, dtype=np.complex128)
    k = len(bits)
    arr = (ctypes.c_int * k)(*[int(b) for b in bits])
    rc = lib.qfo_apply_dense(state.ctypes.data, nbits, mat.ctypes.data, k, arr)
    if rc != 0:
        raise ValueError('qfo_apply_dense failed')


def run_specs(specs, nqubits: int, gate_matrix, state: np.ndarray = None) -> np.ndarray:
    """Circuit.run over (name, params, qubits) specs on a flat vector (qubit q = index bit n-1-q)."""
    if state is None:
        state = np.zeros(1 << nqubits, dtype=np.complex128)
        state[0] = 1.0
    for name, params, qubits in specs:
        apply_dense(state, gate_matrix(name, params), [nqubits - 1 - q for q in qubits])
    return state


def norm2(state: np.ndarray) -> float:
    return float(load().qfo_norm2(state.ctypes.data, int(np.log2(state.size))))


def max_threads() -> int:
    return int(load().qfo_max_threads())


def set_threads(n: int) -> None:
    load().qfo_set_threads(int(n))


def evolve_specs(specs, nqubits: int, gate_matrix, depolarizing_superop) -> np.ndarray:
    """Circuit.evolve over specs on the flat density (qf_oracle.evolve_specs with the C kernel instead of einsum:
    rho is a [2]*(2n) tensor, ket axes first, so ket qubit q is index bit 2n-1-q and bra qubit q index bit n-1-q;
    a gate acts as kron(U, conj U) on (ket qubits, bra qubits), DEPOLARIZING as its Kraus-sum superoperator).
    Starts from |0..0><0..0|. For densities too large for einsum's temporaries (14 qubits = 2^28 elements)."""
    n = nqubits
    rho = np.zeros(1 << (2 * n), dtype=np.complex128)
    rho[0] = 1.0
    for name, params, qubits in specs:
        if name == 'DEPOLARIZING':
            sup = depolarizing_superop(params[0])
        else:
            u = gate_matrix(name, params)
            sup = np.kron(u, u.conj())
        bits = [2 * n - 1 - q for q in qubits] + [n - 1 - q for q in qubits]
        apply_dense(rho, sup, bits)
    return rho

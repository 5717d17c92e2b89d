"""Structural classification of small operators (host side, numpy).

The kernels specialise on operator structure (D = diagonal, C = controlled, P = permutation,
X = dense). Classification is exact (entries compared with 0 and 1 exactly): gate constructors produce exact
zeros and ones, and treating a 1e-17 as zero would silently change results.
"""
from typing import List, Tuple

import numpy as np


def as_matrix(op, k: int = None) -> np.ndarray:
    mat = np.asarray(op, dtype=np.complex128)
    if k is None:
        k = int(round(np.log2(mat.size))) // 2
    return np.ascontiguousarray(mat.reshape(1 << k, 1 << k))


def is_diagonal(mat: np.ndarray) -> bool:
    return bool(np.count_nonzero(mat - np.diag(np.diagonal(mat))) == 0)


def is_identity(mat: np.ndarray) -> bool:
    return bool(np.array_equal(mat, np.eye(mat.shape[0], dtype=mat.dtype)))


def peel_controls(mat: np.ndarray, k: int) -> Tuple[List[int], List[int], np.ndarray]:
    """Split gate qubits into controls and targets.

    Gate qubit q (0 = MSB of the matrix index) is a control when the operator is `P0(q) (x) I + P1(q) (x) U'`,
    i.e. identity whenever q is 0 and never mixing q=0 with q=1. Returns (control qubits, target qubits,
    reduced matrix over the target qubits in their original relative order).
    """
    qubits = list(range(k))
    controls: List[int] = []
    cur = mat
    changed = True
    while changed and len(qubits) > 1:
        changed = False
        kk = len(qubits)
        for pos in range(kk):
            sh = kk - 1 - pos
            idx0 = [i for i in range(1 << kk) if not (i >> sh) & 1]
            idx1 = [i for i in range(1 << kk) if (i >> sh) & 1]
            b00 = cur[np.ix_(idx0, idx0)]
            if (np.count_nonzero(cur[np.ix_(idx0, idx1)]) == 0 and np.count_nonzero(cur[np.ix_(idx1, idx0)]) == 0
                    and is_identity(b00)):
                controls.append(qubits[pos])
                cur = np.ascontiguousarray(cur[np.ix_(idx1, idx1)])
                del qubits[pos]
                changed = True
                break
    return controls, qubits, cur


def g1_kind(m: np.ndarray) -> int:
    """Structure of a 2x2 operator; values match the QFB_G1_* enum in csrc/qfb_plan.h."""
    m = np.asarray(m, dtype=np.complex128).reshape(2, 2)
    if np.array_equal(m, np.array([[0, 1], [1, 0]], dtype=np.complex128)):
        return 3  # SWAPX
    if m[0, 0] == 0 and m[1, 1] == 0:
        return 4  # ANTIDIAG
    if np.count_nonzero(m.imag) == 0:
        return 1  # REAL
    if m[0, 0].imag == 0 and m[1, 1].imag == 0 and m[0, 1].real == 0 and m[1, 0].real == 0:
        return 2  # RXLIKE
    return 0      # GENERAL

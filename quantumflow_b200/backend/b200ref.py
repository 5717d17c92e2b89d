"""The b200 backend as the REFERENCE's own classes see it (the module that the
reference's `quantumflow/backend/__init__.py:14-23` star-imports when QUANTUMFLOW_BACKEND=b200).

The reference builds every state, density, gate and channel through `bk.astensorproduct`
(`quantumflow/qubits.py:103`, `states.py:50,271`, `ops.py:111`) and applies everything with `bk.tensormul`; it has
no notion of "operator tensors on the host". So under this module EVERY product tensor is an amplitude tensor: a
complex128 CUDA tensor, and every `tensormul` / `inner` / `outer` / `conj` / `transpose` / `trace` /
`productdiag` on it is a libqfb200 kernel. Nothing is computed on the CPU: a host tensor that reaches `tensormul`
is uploaded, never multiplied on the host. Small tensors remember the host array they were built from, so that
applying a gate does not read its 2^k x 2^k matrix back from the device (the matrix is a launch parameter).

The mirror package (`quantumflow_b200`) uses `b200bk` instead, which keeps gate / channel operators on the host as
planner inputs; this module is only about running the reference's unmodified classes on the GPU.
"""
import math
import typing

import numpy as np
import torch

from . import b200bk as _bk
from .b200bk import *          # noqa: F401,F403
from .b200bk import __all__ as _bk_all

BACKEND = 'b200'
SEED = None

_HOST_COPY_LIMIT = 1 << 12     # elements: operators up to 6 qubits (4096-element channels)


def _remember(tensor: torch.Tensor, host: np.ndarray) -> torch.Tensor:
    if host.size <= _HOST_COPY_LIMIT:
        tensor._qfb_host = host
    return tensor


def host_value(tensor) -> np.ndarray:
    """Host copy of a (small) operator tensor: the array it was built from when known, else a device read."""
    cached = getattr(tensor, '_qfb_host', None)
    if cached is not None:
        return cached
    return _bk.evaluate(tensor)


def astensor(array) -> torch.Tensor:
    """complex128 CUDA tensor (torch tensors keep their values; host data is uploaded)."""
    if isinstance(array, torch.Tensor):
        t = array if array.dtype == _bk.CTYPE else array.to(_bk.CTYPE)
        return t if t.is_cuda else t.to(_bk.device())
    host = np.array(array, dtype=np.complex128, copy=True)
    return _remember(torch.from_numpy(host).to(_bk.device()), host)


def astensorproduct(array) -> torch.Tensor:
    """astensor + reshape to [2]*n (numpybk.py:105-112); always resident in HBM."""
    tensor = astensor(array)
    n = int(math.log2(tensor.numel())) if tensor.numel() > 0 else 0
    if tensor.numel() != (1 << n):
        raise ValueError('Number of elements is not a power of two')
    out = tensor.reshape([2] * n)
    cached = getattr(tensor, '_qfb_host', None)
    if cached is not None:
        out._qfb_host = cached.reshape([2] * n)
    return out


def tensormul(tensor0, tensor1, indices: typing.List[int]) -> torch.Tensor:
    """numpybk.py:159-214 semantics on amplitude tensors: tensor1 lives in HBM (a host tensor is uploaded, never
    multiplied on the host), tensor0's 2^k x 2^k matrix is a launch parameter."""
    n = _bk.rank(tensor1)
    k = _bk.rank(tensor0) // 2
    indices = [int(i) for i in indices]
    assert k == len(indices)
    if len(set(indices)) != k or any(i < 0 or i >= n for i in indices):
        raise ValueError('tensormul: bad indices {}'.format(indices))
    if not _bk.is_amplitudes(tensor1):
        tensor1 = astensorproduct(tensor1)
    from .. import engine
    mat = np.ascontiguousarray(host_value(tensor0), dtype=np.complex128).reshape(1 << k, 1 << k)
    bits = [n - 1 - i for i in indices]
    return engine.apply_operator(tensor1.contiguous(), mat, bits)


def inner(tensor0, tensor1):
    return _bk.inner(astensor(tensor0), astensor(tensor1))


def outer(tensor0, tensor1):
    return _bk.outer(astensor(tensor0), astensor(tensor1))


__all__ = list(_bk_all) + ['BACKEND', 'SEED', 'host_value']

"""QuantumFlow tensor backend for B200: torch tensors as containers, hand-written sm_100a kernels as arithmetic.

This module exports the names of the reference's backend contract (quantumflow/backend/__init__.py:25-35 plus
`size`, tests/test_backend.py:108-114). Two tensor domains exist:

* amplitude tensors -- states / densities, `torch.complex128` on a CUDA device, shape [2]*n. Every operation on
  them (tensormul, inner, outer, conj, productdiag, trace, transpose ...) is a kernel of libqfb200.so reached
  through the C ABI. There is no CPU implementation for this domain; without a GPU these calls raise.
* operator tensors -- gate / channel operators (at most a few thousand elements), `torch.complex128` on the
  host. They are planner inputs: composing them (`Gate @ Gate`, `aschannel`, `H`) is host-side matrix algebra,
  stays differentiable for the QAOA bridge, and they enter kernels as launch parameters.

`tensormul(op, amplitudes, indices)` is the hot call (reference: numpybk.py:159-214).
"""
import math
import string
import typing

import numpy as np
import torch

from .. import _lib
from .. import config as _config

TL = torch
name = TL.__name__
version = TL.__version__

CTYPE = torch.complex128
FTYPE = torch.float64
TENSOR = torch.Tensor
BKTensor = typing.Any
TensorLike = typing.Any

# 16*2^n bytes per state: n=33 is 128 GiB and fits one 180 GB B200; beyond that the state is sharded
MAX_QUBITS = 62

EINSUM_SUBSCRIPTS = string.ascii_lowercase + string.ascii_uppercase

DEVICE = 'gpu'

# largest host tensor that tensormul treats as an operator (a 6-qubit channel has 4^6 = 4096 elements ... 2^16 covers
# 8-qubit gates); anything larger on the host is an amplitude tensor in the wrong place
HOST_OPERATOR_LIMIT = 1 << 16
_STAGED_MIN_BYTES = 1 << 28            # host <-> device copies of at least 256 MiB go through engine.upload / download

from math import pi  # noqa: E402,F401  (the reference backends re-export pi)


def gpu_available() -> bool:
    return torch.cuda.is_available()


def device() -> torch.device:
    """CUDA device that holds amplitude tensors."""
    if not torch.cuda.is_available():
        raise RuntimeError('quantumflow_b200: no CUDA device; states and densities live in HBM and there is '
                           'no CPU fallback')
    idx = _config.DEVICE_INDEX if _config.DEVICE_INDEX is not None else torch.cuda.current_device()
    return torch.device('cuda', idx)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def is_amplitudes(tensor) -> bool:
    return isinstance(tensor, torch.Tensor) and tensor.is_cuda


# ---------------------------------------------------------------------------------------------------------
# conversion
# ---------------------------------------------------------------------------------------------------------

def _stack_nested(obj):
    """Pack a nested list whose leaves may be autograd tensors, without detaching."""
    if isinstance(obj, torch.Tensor):
        return obj.to(CTYPE)
    if isinstance(obj, (list, tuple)):
        return torch.stack([_stack_nested(o) for o in obj])
    return torch.tensor(complex(obj), dtype=CTYPE)


def _has_tensor_leaf(obj) -> bool:
    if isinstance(obj, torch.Tensor):
        return True
    if isinstance(obj, (list, tuple)):
        return any(_has_tensor_leaf(o) for o in obj)
    return False


def astensor(array: TensorLike) -> BKTensor:
    """Convert to a complex128 torch tensor. Torch inputs keep their device; anything else lands on the host."""
    if isinstance(array, torch.Tensor):
        return array if array.dtype == CTYPE else array.to(CTYPE)
    if isinstance(array, (list, tuple)) and _has_tensor_leaf(array):
        return _stack_nested(array)
    return torch.from_numpy(np.array(array, dtype=np.complex128, copy=True))


def astensorproduct(array: TensorLike) -> BKTensor:
    """astensor + reshape to [2]*n (numpybk.py:105-112)."""
    tensor = astensor(array)
    n = int(math.log2(tensor.numel())) if tensor.numel() > 0 else 0
    if tensor.numel() != (1 << n):
        raise ValueError('Number of elements is not a power of two')
    return tensor.reshape([2] * n)


def asamplitudes(array: TensorLike) -> BKTensor:
    """Product tensor resident in HBM (contiguous complex128 on the CUDA device)."""
    if isinstance(array, np.ndarray) and array.nbytes >= _STAGED_MIN_BYTES and array.size == 1 << int(
            math.log2(array.size)):
        # a state-sized host array: pieces through pinned staging buffers instead of one pageable copy
        from .. import engine
        n = int(math.log2(array.size))
        return engine.upload(array, device()).reshape([2] * n)
    tensor = astensorproduct(array)
    if not tensor.is_cuda:
        tensor = tensor.to(device())
    return tensor.contiguous()


def evaluate(tensor: BKTensor) -> TensorLike:
    """Value of a tensor as a numpy array (device -> host copy for amplitude tensors)."""
    if isinstance(tensor, torch.Tensor):
        if tensor.is_cuda and tensor.dtype == CTYPE and tensor.numel() * 16 >= _STAGED_MIN_BYTES \
                and tensor.is_contiguous():
            from .. import engine
            return engine.download(tensor).reshape(tuple(tensor.shape))
        return tensor.detach().cpu().numpy()
    return np.asarray(tensor)


def rank(tensor: BKTensor) -> int:
    return len(tensor.shape)


def size(tensor: BKTensor) -> int:
    return int(np.prod(tensor.shape)) if len(tensor.shape) else 1


def ccast(value) -> TensorLike:
    if isinstance(value, torch.Tensor):
        return value.to(CTYPE)
    return complex(value)


def fcast(value) -> TensorLike:
    if isinstance(value, torch.Tensor):
        return value.to(FTYPE)
    return float(value)


def set_random_seed(seed: int) -> None:
    """The shared RNG stream is numpy's global RandomState (sampling happens with numpy calls)."""
    np.random.seed(seed)
    torch.manual_seed(seed)


def getitem(tensor: BKTensor, key: typing.Any) -> BKTensor:
    return tensor.__getitem__(key)


# ---------------------------------------------------------------------------------------------------------
# scalar / elementwise math (operator construction; accepts python scalars, numpy values and torch tensors)
# ---------------------------------------------------------------------------------------------------------

def _is_t(x) -> bool:
    return isinstance(x, torch.Tensor)


def _unary(tfn, nfn):
    def fn(x):
        if _is_t(x):
            return tfn(x)
        return nfn(x)
    return fn


sqrt = _unary(torch.sqrt, np.sqrt)
exp = _unary(torch.exp, np.exp)
cos = _unary(torch.cos, np.cos)
sin = _unary(torch.sin, np.sin)
arccos = _unary(torch.arccos, np.arccos)
real = _unary(torch.real, np.real)
imag = _unary(torch.imag, np.imag)
absolute = _unary(torch.abs, np.absolute)


def cis(theta) -> BKTensor:
    r"""cis(theta) = exp(i theta)"""
    if _is_t(theta):
        return torch.exp(theta.to(CTYPE) * 1.0j)
    return np.exp(theta * 1.0j)


def minimum(t0, t1):
    if _is_t(t0) or _is_t(t1):
        t0 = t0 if _is_t(t0) else torch.as_tensor(t0)
        t1 = t1 if _is_t(t1) else torch.as_tensor(t1, dtype=t0.dtype, device=t0.device)
        return torch.minimum(t0, t1.to(t0.device))
    return np.minimum(t0, t1)


def sum(tensor, axis=None, keepdims=False):  # noqa: A001  (name fixed by the contract)
    if _is_t(tensor):
        if axis is None:
            return torch.sum(tensor)
        return torch.sum(tensor, dim=axis, keepdim=keepdims)
    return np.sum(tensor, axis=axis, keepdims=keepdims)


def matmul(t0, t1):
    return torch.matmul(astensor(t0), astensor(t1))


def diag(tensor):
    return torch.diag(tensor) if _is_t(tensor) else np.diag(tensor)


def einsum(subscripts, *operands):
    return torch.einsum(subscripts, *[astensor(o) for o in operands])


def reshape(tensor: BKTensor, shape) -> BKTensor:
    return tensor.reshape(tuple(int(s) for s in shape))


def conj(tensor: BKTensor) -> BKTensor:
    if is_amplitudes(tensor) and not tensor.requires_grad:
        lib = _lib.load()
        src = tensor.contiguous()
        out = torch.empty_like(src)
        _lib.check(lib.qfb_conj(out.data_ptr(), src.data_ptr(), src.numel(), _stream()))
        return out
    if _is_t(tensor):
        return torch.conj(tensor).resolve_conj()
    return np.conj(tensor)


def transpose(tensor: BKTensor, perm=None) -> BKTensor:
    """Axis permutation. On a [2]*n amplitude tensor this is a bit-permutation sweep (qfb_permute_bits)."""
    nd = len(tensor.shape)
    if perm is None:
        perm = list(range(nd))[::-1]
    perm = [int(p) for p in perm]
    if is_amplitudes(tensor) and not tensor.requires_grad and all(s == 2 for s in tensor.shape) and nd > 0:
        lib = _lib.load()
        src = tensor.contiguous()
        out = torch.empty_like(src)
        # output axis j reads input axis perm[j]; axis a <-> flat bit nd-1-a
        bitperm = [0] * nd
        for j, p in enumerate(perm):
            bitperm[nd - 1 - j] = nd - 1 - p
        _lib.check(lib.qfb_permute_bits(out.data_ptr(), src.data_ptr(), nd, _lib.int_array(bitperm), 0,
                                        _stream()))
        return out
    if _is_t(tensor):
        return tensor.permute(perm).contiguous()
    return np.transpose(tensor, perm)


def trace(tensor: BKTensor) -> BKTensor:
    """Trace of a square matrix (qubits.py:185-197 reshapes to [2^m, 2^m] first)."""
    if is_amplitudes(tensor) and not tensor.requires_grad and len(tensor.shape) == 2:
        lib = _lib.load()
        dim = tensor.shape[0]
        nq = int(math.log2(dim))
        if (1 << nq) == dim and tensor.shape[1] == dim:
            src = tensor.contiguous()
            out = torch.empty(2, dtype=FTYPE, device=src.device)
            _lib.check(lib.qfb_density_trace(src.data_ptr(), nq, out.data_ptr(), _stream()))
            return torch.view_as_complex(out)
    if _is_t(tensor):
        return torch.trace(tensor)
    return np.trace(tensor)


def productdiag(tensor: BKTensor) -> BKTensor:
    """Matrix diagonal of a [2]*2n product tensor as a [2]*n tensor (numpybk.py:150-156)."""
    nd = rank(tensor)
    n = nd // 2
    if is_amplitudes(tensor) and not tensor.requires_grad:
        lib = _lib.load()
        src = tensor.contiguous()
        out = torch.empty([2] * n, dtype=CTYPE, device=src.device)
        _lib.check(lib.qfb_density_diag(src.data_ptr(), n, out.data_ptr(), _stream()))
        return out
    mat = astensor(tensor).reshape(1 << n, 1 << n)
    return torch.diagonal(mat).reshape([2] * n)


# ---------------------------------------------------------------------------------------------------------
# products
# ---------------------------------------------------------------------------------------------------------

def inner(tensor0: BKTensor, tensor1: BKTensor) -> BKTensor:
    """<t0|t1> with the first argument conjugated (np.vdot semantics, numpybk.py:125-128)."""
    if is_amplitudes(tensor0) or is_amplitudes(tensor1):
        from ..autograd import needs_grad, inner_autograd
        t0 = asamplitudes(tensor0) if not is_amplitudes(tensor0) else tensor0.contiguous()
        t1 = asamplitudes(tensor1) if not is_amplitudes(tensor1) else tensor1.contiguous()
        if t0.numel() != t1.numel():
            raise ValueError('inner: size mismatch')
        if needs_grad(t0) or needs_grad(t1):
            return inner_autograd(t0, t1)
        lib = _lib.load()
        out = torch.empty(2, dtype=FTYPE, device=t0.device)
        _lib.check(lib.qfb_vdot(t0.data_ptr(), t1.data_ptr(), t0.numel(), out.data_ptr(), _stream()))
        return torch.view_as_complex(out)
    t0 = astensor(tensor0).reshape(-1)
    t1 = astensor(tensor1).reshape(-1)
    return torch.vdot(t0, t1)


def outer(tensor0: BKTensor, tensor1: BKTensor) -> BKTensor:
    """numpy.outer semantics: both inputs are flattened, result is [size0, size1]."""
    if is_amplitudes(tensor0) or is_amplitudes(tensor1):
        lib = _lib.load()
        t0 = asamplitudes(tensor0) if not is_amplitudes(tensor0) else tensor0.contiguous()
        t1 = asamplitudes(tensor1) if not is_amplitudes(tensor1) else tensor1.contiguous()
        out = torch.empty((t0.numel(), t1.numel()), dtype=CTYPE, device=t0.device)
        _lib.check(lib.qfb_outer(out.data_ptr(), t0.data_ptr(), t0.numel(), t1.data_ptr(), t1.numel(), 0,
                                 _stream()))
        return out
    return torch.outer(astensor(tensor0).reshape(-1), astensor(tensor1).reshape(-1))


def _host_tensormul(t0: torch.Tensor, t1: torch.Tensor, indices) -> torch.Tensor:
    """Operator (x) operator composition on the host (planner-side algebra; differentiable)."""
    n = t1.dim()
    k = len(indices)
    rest = [ax for ax in range(n) if ax not in indices]
    moved = t1.permute(list(indices) + rest).reshape(1 << k, -1)
    res = t0.reshape(1 << k, 1 << k) @ moved
    res = res.reshape([2] * n)
    inverse = [0] * n
    for pos, ax in enumerate(list(indices) + rest):
        inverse[ax] = pos
    return res.permute(inverse).contiguous()


def tensormul(tensor0: BKTensor, tensor1: BKTensor, indices: typing.List[int]) -> BKTensor:
    """out[.. a ..] = sum_in G[a, in] * T[.. in ..]; gate qubit j acts on axis indices[j] of tensor1, the output
    keeps tensor1's axis order (reference: numpybk.py:159-214)."""
    n = rank(tensor1)
    k = rank(tensor0) // 2
    indices = [int(i) for i in indices]
    assert k == len(indices)
    if len(set(indices)) != k or any(i < 0 or i >= n for i in indices):
        raise ValueError('tensormul: bad indices {}'.format(indices))

    if not is_amplitudes(tensor1):
        # operator (x) operator algebra of the planner (Gate @ Gate, aschannel): at most a few thousand elements.
        # A state-sized tensor on the host is a usage error: amplitudes live in HBM and there is no CPU path.
        if size(tensor1) > HOST_OPERATOR_LIMIT:
            raise RuntimeError('tensormul: a {}-element tensor on the host is not an operator; states and densities '
                               'must be amplitude tensors (bk.asamplitudes / qf.State) -- there is no CPU path'
                               .format(size(tensor1)))
        return _host_tensormul(astensor(tensor0), astensor(tensor1), indices)

    from ..autograd import needs_grad, tensormul_autograd
    if needs_grad(tensor0) or needs_grad(tensor1):
        return tensormul_autograd(astensor(tensor0), tensor1, indices)

    from .. import engine
    mat = evaluate(tensor0).reshape(1 << k, 1 << k)
    bits = [n - 1 - i for i in indices]
    return engine.apply_operator(tensor1, mat, bits)


__all__ = [  # noqa: F822
    'BKTensor', 'CTYPE', 'DEVICE', 'FTYPE', 'MAX_QUBITS', 'TENSOR', 'TL', 'TensorLike', 'absolute', 'arccos',
    'astensor', 'ccast', 'cis', 'conj', 'cos', 'diag', 'evaluate', 'exp', 'fcast', 'gpu_available', 'imag',
    'inner', 'minimum', 'outer', 'matmul', 'rank', 'real', 'reshape', 'set_random_seed', 'sin', 'sqrt', 'sum',
    'tensormul', 'trace', 'transpose', 'getitem', 'astensorproduct', 'productdiag', 'EINSUM_SUBSCRIPTS',
    'einsum', 'size', 'asamplitudes', 'is_amplitudes', 'device', 'name', 'version', 'pi']

"""Thin host wrappers over the C ABI for amplitude tensors (torch complex128 CUDA tensors).

Each function launches kernels of libqfb200.so on torch's current stream and returns torch tensors; scalar
results stay on the device unless the caller asks for a Python value. Nothing here computes amplitudes on the
CPU: a missing library or GPU raises.
"""
import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import classify

CTYPE = torch.complex128
FTYPE = torch.float64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_amplitudes(tensor: torch.Tensor) -> torch.Tensor:
    if not (isinstance(tensor, torch.Tensor) and tensor.is_cuda and tensor.dtype == CTYPE):
        raise TypeError('expected a complex128 CUDA tensor (amplitude tensor)')
    return tensor if tensor.is_contiguous() else tensor.contiguous()


def nbits_of(tensor: torch.Tensor) -> int:
    n = tensor.numel()
    nb = int(math.log2(n)) if n > 0 else 0
    if (1 << nb) != n:
        raise ValueError('amplitude tensor size {} is not a power of two'.format(n))
    return nb


def _dptr(arr: np.ndarray):
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def launch_count() -> int:
    return int(_lib.load().qfb_launch_count())


# ---------------------------------------------------------------------------------------------------------
# host <-> device transfer of large amplitude arrays
# ---------------------------------------------------------------------------------------------------------
# A pageable 16 GiB numpy array moves at a few GB/s through torch's default path (one staged copy at a time, the
# reference-shaped call State(array) -> Circuit.run -> asarray spent 12 of its 13 seconds there). These helpers
# cut the array into pieces that go through two pinned staging buffers: the (multi-threaded) host copy of piece
# i+1 runs while piece i crosses PCIe.

STAGED_TRANSFER_MIN_BYTES = 1 << 28       # arrays of at least 256 MiB take the staged path
_STAGE_BYTES = 1 << 28
_stage_buffers = {}


def _staging(device: torch.device):
    key = (device.index,)
    if key not in _stage_buffers:
        _stage_buffers[key] = [torch.empty(_STAGE_BYTES // 16, dtype=CTYPE).pin_memory() for _ in range(2)]
    return _stage_buffers[key]


def upload(array: np.ndarray, device: torch.device) -> torch.Tensor:
    """complex128 host array -> flat CUDA tensor through the pinned staging ring."""
    src = torch.from_numpy(np.ascontiguousarray(array, dtype=np.complex128).reshape(-1))
    out = torch.empty(src.numel(), dtype=CTYPE, device=device)
    stage = _staging(device)
    piece = stage[0].numel()
    stream = torch.cuda.Stream(device)
    # `out` may be a recycled block that work already queued on the current stream still reads: stay behind it
    stream.wait_stream(torch.cuda.current_stream(device))
    done = [None, None]
    for i, lo in enumerate(range(0, src.numel(), piece)):
        n = min(piece, src.numel() - lo)
        buf = stage[i % 2]
        if done[i % 2] is not None:
            done[i % 2].synchronize()                 # the previous copy out of this buffer has finished
        buf[:n].copy_(src[lo:lo + n])                 # host -> pinned (parallel memcpy)
        with torch.cuda.stream(stream):
            out[lo:lo + n].copy_(buf[:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            done[i % 2] = ev
    torch.cuda.current_stream(device).wait_stream(stream)
    stream.synchronize()
    return out


def download(tensor: torch.Tensor) -> np.ndarray:
    """Flat complex128 CUDA tensor -> numpy array through the pinned staging ring."""
    flat = tensor.detach().reshape(-1)
    out = np.empty(flat.numel(), dtype=np.complex128)
    dst = torch.from_numpy(out)
    device = flat.device
    stage = _staging(device)
    piece = stage[0].numel()
    stream = torch.cuda.Stream(device)
    stream.wait_stream(torch.cuda.current_stream(device))
    pending = []
    for i, lo in enumerate(range(0, flat.numel(), piece)):
        n = min(piece, flat.numel() - lo)
        buf = stage[i % 2]
        if len(pending) == 2:                          # this buffer's previous piece must reach `out` first
            ev, plo, pn, pbuf = pending.pop(0)
            ev.synchronize()
            dst[plo:plo + pn].copy_(pbuf[:pn])
        with torch.cuda.stream(stream):
            buf[:n].copy_(flat[lo:lo + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        pending.append((ev, lo, n, buf))
    for ev, plo, pn, pbuf in pending:
        ev.synchronize()
        dst[plo:plo + pn].copy_(pbuf[:pn])
    return out


# ---------------------------------------------------------------------------------------------------------
# gate application
# ---------------------------------------------------------------------------------------------------------

def apply_operator(tensor: torch.Tensor, mat: np.ndarray, bits: Sequence[int], inplace: bool = False,
                   index_hi: int = 0, classify_structure: bool = True) -> torch.Tensor:
    """Apply a k-bit operator (row-major 2^k x 2^k, gate qubit 0 = MSB) to index bits `bits` of `tensor`.

    Structure is exploited: diagonal operators go to the diagonal kernel, control qubits are peeled off so that
    only the control=1 groups are touched.
    """
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    nb = nbits_of(src)
    k = len(bits)
    mat = classify.as_matrix(mat, k)
    if inplace and src.data_ptr() != tensor.data_ptr():
        raise ValueError('in-place application needs a contiguous tensor')
    dst = src if inplace else torch.empty_like(src)
    st = _stream()
    bits = [int(b) for b in bits]
    if classify_structure and k >= 1 and classify.is_diagonal(mat):
        table = np.ascontiguousarray(np.diagonal(mat)).view(np.float64)
        _lib.check(lib.qfb_apply_diag(dst.data_ptr(), src.data_ptr(), nb, _dptr(table), k, _lib.int_array(bits),
                                      index_hi, st))
        return dst
    ctrl_bits = []
    tbits = bits
    if classify_structure and k >= 2:
        controls, targets, reduced = classify.peel_controls(mat, k)
        if controls:
            ctrl_bits = [bits[q] for q in controls]
            tbits = [bits[q] for q in targets]
            mat = reduced
    flat = np.ascontiguousarray(mat).view(np.float64)
    _lib.check(lib.qfb_apply_dense(dst.data_ptr(), src.data_ptr(), nb, _dptr(flat), len(tbits),
                                   _lib.int_array(tbits), len(ctrl_bits), _lib.int_array(ctrl_bits), index_hi, st))
    return dst


# ---------------------------------------------------------------------------------------------------------
# reductions / read-out
# ---------------------------------------------------------------------------------------------------------

def norm2(tensor: torch.Tensor) -> torch.Tensor:
    """sum |a|^2 as a 0-d float64 device tensor."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    out = torch.empty(1, dtype=FTYPE, device=src.device)
    _lib.check(lib.qfb_norm2(src.data_ptr(), src.numel(), out.data_ptr(), _stream()))
    return out.reshape(())


def vdot(t0: torch.Tensor, t1: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    a = _require_amplitudes(t0)
    b = _require_amplitudes(t1)
    if a.numel() != b.numel():
        raise ValueError('vdot: size mismatch')
    out = torch.empty(2, dtype=FTYPE, device=a.device)
    _lib.check(lib.qfb_vdot(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream()))
    return torch.view_as_complex(out)


def probabilities(tensor: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    out = torch.empty(src.shape, dtype=FTYPE, device=src.device)
    _lib.check(lib.qfb_probs(src.data_ptr(), src.numel(), out.data_ptr(), _stream()))
    return out


def expectation_diag(tensor: torch.Tensor, diag: torch.Tensor) -> torch.Tensor:
    """sum_i diag[i] |a_i|^2 ; `diag` is a float64 device tensor with as many elements as `tensor`."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    if not (diag.is_cuda and diag.dtype == FTYPE and diag.numel() == src.numel()):
        raise TypeError('expectation_diag: diag must be a float64 CUDA tensor of matching size')
    diag = diag.contiguous()
    out = torch.empty(1, dtype=FTYPE, device=src.device)
    _lib.check(lib.qfb_expect_diag(src.data_ptr(), diag.data_ptr(), src.numel(), out.data_ptr(), _stream()))
    return out.reshape(())


def marginal(tensor: torch.Tensor, bit: int) -> torch.Tensor:
    """[p(bit=0), p(bit=1)] (unnormalised) as a float64[2] device tensor."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    out = torch.empty(2, dtype=FTYPE, device=src.device)
    _lib.check(lib.qfb_marginal(src.data_ptr(), nbits_of(src), int(bit), out.data_ptr(), _stream()))
    return out


def collapse(tensor: torch.Tensor, bit: int, value: int, scale: float, inplace: bool = False) -> torch.Tensor:
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    dst = src if inplace else torch.empty_like(src)
    _lib.check(lib.qfb_collapse(dst.data_ptr(), src.data_ptr(), nbits_of(src), int(bit), int(value), float(scale),
                                _stream()))
    return dst


def scale(tensor: torch.Tensor, factor: complex, inplace: bool = False) -> torch.Tensor:
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    dst = src if inplace else torch.empty_like(src)
    factor = complex(factor)
    _lib.check(lib.qfb_scale(dst.data_ptr(), src.data_ptr(), src.numel(), factor.real, factor.imag, _stream()))
    return dst


def normalize_by_norm2(tensor: torch.Tensor, n2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tensor / sqrt(<t|t>) with the norm kept on the device (State.normalize, states.py:108-111)."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    if n2 is None:
        n2 = norm2(src)
    n2 = n2.reshape(1).contiguous()
    dst = torch.empty_like(src)
    _lib.check(lib.qfb_scale_rsqrt_dev(dst.data_ptr(), src.data_ptr(), src.numel(), n2.data_ptr(), _stream()))
    return dst


def divide_by_device_scalar(tensor: torch.Tensor, scalar: torch.Tensor) -> torch.Tensor:
    """tensor / scalar for a complex 0-d device scalar (Density.normalize divides by the trace)."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    sc = torch.view_as_real(scalar.to(CTYPE).reshape(1)).contiguous()
    dst = torch.empty_like(src)
    _lib.check(lib.qfb_scale_cdiv_dev(dst.data_ptr(), src.data_ptr(), src.numel(), sc.data_ptr(), _stream()))
    return dst


def axpby(a: torch.Tensor, alpha: complex, b: Optional[torch.Tensor] = None, beta: complex = 0.0) -> torch.Tensor:
    lib = _lib.load()
    a = _require_amplitudes(a)
    alpha = complex(alpha)
    beta = complex(beta)
    bptr = None
    if b is not None:
        b = _require_amplitudes(b)
        if b.numel() != a.numel():
            raise ValueError('axpby: size mismatch')
        bptr = b.data_ptr()
    dst = torch.empty_like(a)
    _lib.check(lib.qfb_axpby(dst.data_ptr(), a.data_ptr(), alpha.real, alpha.imag, bptr, beta.real, beta.imag,
                             a.numel(), _stream()))
    return dst


def outer(t0: torch.Tensor, t1: torch.Tensor, conj_second: bool = False) -> torch.Tensor:
    lib = _lib.load()
    a = _require_amplitudes(t0)
    b = _require_amplitudes(t1)
    out = torch.empty(a.numel() * b.numel(), dtype=CTYPE, device=a.device)
    _lib.check(lib.qfb_outer(out.data_ptr(), a.data_ptr(), a.numel(), b.data_ptr(), b.numel(),
                             1 if conj_second else 0, _stream()))
    return out


def permute_bits(tensor: torch.Tensor, perm: Sequence[int], conj: bool = False) -> torch.Tensor:
    """dst index bit j <- src index bit perm[j]."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    nb = nbits_of(src)
    if len(perm) != nb:
        raise ValueError('permute_bits: need {} entries'.format(nb))
    dst = torch.empty_like(src)
    _lib.check(lib.qfb_permute_bits(dst.data_ptr(), src.data_ptr(), nb, _lib.int_array(perm), 1 if conj else 0,
                                    _stream()))
    return dst


def density_diag(tensor: torch.Tensor, nq: int) -> torch.Tensor:
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    out = torch.empty([2] * nq, dtype=CTYPE, device=src.device)
    _lib.check(lib.qfb_density_diag(src.data_ptr(), int(nq), out.data_ptr(), _stream()))
    return out


def density_trace(tensor: torch.Tensor, nq: int) -> torch.Tensor:
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    out = torch.empty(2, dtype=FTYPE, device=src.device)
    _lib.check(lib.qfb_density_trace(src.data_ptr(), int(nq), out.data_ptr(), _stream()))
    return torch.view_as_complex(out).reshape(())


def partial_trace(tensor: torch.Tensor, keep_bits: Sequence[int], trace_masks: Sequence[int]) -> torch.Tensor:
    """Sum over the settings in which all copies of each traced qubit agree (qfb_partial_trace): the result has
    len(keep_bits) index bits, output bit b <- input bit keep_bits[b]; trace_masks[t] = OR of the input bits of
    traced qubit t. Returns a flat tensor of 2^len(keep_bits) amplitudes."""
    lib = _lib.load()
    src = _require_amplitudes(tensor)
    nb = nbits_of(src)
    out = torch.empty(1 << len(keep_bits), dtype=CTYPE, device=src.device)
    masks = (ctypes.c_uint64 * max(1, len(trace_masks)))(*[int(m) for m in trace_masks])
    _lib.check(lib.qfb_partial_trace(out.data_ptr(), src.data_ptr(), nb, len(keep_bits), _lib.int_array(keep_bits),
                                     len(trace_masks), masks, _stream()))
    return out


def sample_search(probs: torch.Tensor, uniforms: np.ndarray) -> np.ndarray:
    """Indices i with cdf(i-1) <= u*total < cdf(i) for each u (device block sums + host chunk search)."""
    lib = _lib.load()
    if not (probs.is_cuda and probs.dtype == FTYPE):
        raise TypeError('sample_search: probs must be a float64 CUDA tensor')
    probs = probs.contiguous()
    u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1)
    out = np.zeros(u.size, dtype=np.uint64)
    _lib.check(lib.qfb_sample_search(probs.data_ptr(), probs.numel(), _dptr(u), int(u.size),
                                     out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), _stream()))
    return out


def gate_grad(grad_out: torch.Tensor, psi: torch.Tensor, bits: Sequence[int]) -> torch.Tensor:
    """grad_M[r][c] = sum_groups g[base|off[r]] * conj(psi[base|off[c]]) as a host complex128 [2^k, 2^k] tensor."""
    lib = _lib.load()
    g = _require_amplitudes(grad_out)
    p = _require_amplitudes(psi)
    k = len(bits)
    out = torch.empty((1 << k, 1 << k), dtype=CTYPE, device=g.device)
    _lib.check(lib.qfb_gate_grad(g.data_ptr(), p.data_ptr(), nbits_of(p), k, _lib.int_array(bits), out.data_ptr(),
                                 _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------
# plans
# ---------------------------------------------------------------------------------------------------------

class UploadedPlan:
    """A binary plan resident in device memory; replayable on any state with the same number of index bits."""

    def __init__(self, blob: bytes) -> None:
        lib = _lib.load()
        self._lib = lib
        self._blob = blob
        handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        _lib.check(lib.qfb_plan_upload(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(handle),
                                       _stream()))
        self._handle = handle

    def launch(self, tensor: torch.Tensor, index_hi: int = 0) -> None:
        """Execute in place on `tensor`."""
        src = _require_amplitudes(tensor)
        if src.data_ptr() != tensor.data_ptr():
            raise ValueError('plan execution is in place and needs a contiguous tensor')
        _lib.check(self._lib.qfb_plan_launch(self._handle, src.data_ptr(), nbits_of(src), int(index_hi),
                                             _stream()))

    def _info(self, sweep: int = 0):
        n, mask, spec = ctypes.c_int(0), ctypes.c_uint64(0), ctypes.c_int(0)
        _lib.check(self._lib.qfb_plan_sweep_info(self._handle, int(sweep), ctypes.byref(n), ctypes.byref(mask),
                                                 ctypes.byref(spec)))
        return n.value, mask.value, bool(spec.value)

    @property
    def nsweeps(self) -> int:
        return self._info()[0]

    @property
    def specialised(self) -> bool:
        """Does the plan run on sweep-specialised kernels (csrc/qfb_jit.cu)? Only those can be launched on slices."""
        return self._info()[2]

    def nontile_mask(self, sweep: int) -> int:
        """Index bits outside the tile of sweep `sweep`: amplitudes that differ in one of them never meet in it."""
        return self._info(sweep)[1]

    def launch_part(self, tensor: torch.Tensor, first_sweep: int, nsweeps: int, index_hi: int = 0,
                    fix_mask: int = 0, fix_value: int = 0, ctas_per_sm: int = 0) -> None:
        """Sweeps [first_sweep, first_sweep + nsweeps) in place on `tensor`; with fix_mask on the slice of the state
        whose index bits fix_mask equal fix_value (bits of nontile_mask of every launched sweep)."""
        src = _require_amplitudes(tensor)
        if src.data_ptr() != tensor.data_ptr():
            raise ValueError('plan execution is in place and needs a contiguous tensor')
        _lib.check(self._lib.qfb_plan_launch_part(self._handle, src.data_ptr(), nbits_of(src), int(index_hi),
                                                  int(first_sweep), int(nsweeps), int(fix_mask), int(fix_value),
                                                  int(ctas_per_sm), _stream()))

    def close(self) -> None:
        if self._handle is not None and self._handle.value:
            self._lib.qfb_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

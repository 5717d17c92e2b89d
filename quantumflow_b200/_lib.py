"""ctypes binding of libqfb200.so -- the only door between the Python host code and the CUDA kernels.

Every prototype here mirrors a declaration in include/qfb200.h. There is no alternative implementation:
if the library cannot be loaded, or a call fails, an exception is raised (no CPU fallback).
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER

from . import _build

_c_int_p = POINTER(c_int)
_c_double_p = POINTER(c_double)

# name -> (restype, argtypes); kept as data so tests can check the export list against include/qfb200.h
PROTOTYPES = {
    'qfb_version': (c_int, []),
    'qfb_last_error': (c_char_p, []),
    'qfb_device_props': (c_int, [c_int, _c_int_p, _c_int_p, _c_int_p, POINTER(c_size_t), POINTER(c_size_t)]),
    'qfb_apply_dense': (c_int, [c_void_p, c_void_p, c_int, _c_double_p, c_int, _c_int_p, c_int, _c_int_p,
                                c_uint64, c_void_p]),
    'qfb_apply_diag': (c_int, [c_void_p, c_void_p, c_int, _c_double_p, c_int, _c_int_p, c_uint64, c_void_p]),
    'qfb_run_plan': (c_int, [c_void_p, c_int, c_uint64, c_void_p, c_size_t, c_void_p]),
    'qfb_plan_validate': (c_int, [c_void_p, c_size_t]),
    'qfb_plan_upload': (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_void_p]),
    'qfb_plan_launch': (c_int, [c_void_p, c_void_p, c_int, c_uint64, c_void_p]),
    'qfb_plan_destroy': (c_int, [c_void_p]),
    'qfb_plan_launch_part': (c_int, [c_void_p, c_void_p, c_int, c_uint64, c_int, c_int, c_uint64, c_uint64, c_int,
                                     c_void_p]),
    'qfb_plan_sweep_info': (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_uint64), POINTER(c_int)]),
    'qfb_jit_ptx': (c_int, [c_void_p, c_size_t, c_int, c_char_p, c_size_t, POINTER(c_size_t), POINTER(c_size_t)]),
    'qfb_jit_check': (c_int, [c_void_p, c_size_t, c_char_p, c_size_t]),
    'qfb_jit_source': (c_int, [c_void_p, c_size_t, c_int, c_uint64, c_char_p, c_size_t, POINTER(c_size_t), c_void_p,
                               c_size_t, POINTER(c_size_t), _c_int_p, POINTER(c_size_t), _c_int_p]),
    'qfb_small_circuit_run': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'qfb_small_circuit_adjoint': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    'qfb_small_circuit_scratch_doubles': (c_size_t, [c_int, c_int]),
    'qfb_launch_count': (c_uint64, []),
    'qfb_vdot': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_norm2': (c_int, [c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_probs': (c_int, [c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_expect_diag': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_marginal': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'qfb_collapse': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p]),
    'qfb_scale': (c_int, [c_void_p, c_void_p, c_uint64, c_double, c_double, c_void_p]),
    'qfb_scale_rsqrt_dev': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_scale_cdiv_dev': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    'qfb_axpby': (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_double, c_double, c_uint64,
                          c_void_p]),
    'qfb_outer': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_int, c_void_p]),
    'qfb_conj': (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    'qfb_density_diag': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'qfb_density_trace': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'qfb_partial_trace': (c_int, [c_void_p, c_void_p, c_int, c_int, _c_int_p, c_int, POINTER(c_uint64), c_void_p]),
    'qfb_permute_bits': (c_int, [c_void_p, c_void_p, c_int, _c_int_p, c_int, c_void_p]),
    'qfb_sample_search': (c_int, [c_void_p, c_uint64, _c_double_p, c_int, POINTER(c_uint64), c_void_p]),
    'qfb_plan_count_executed': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_uint64, c_uint64, c_double,
                                        c_int64, _c_int_p]),
    'qfb_plan_refine_tile': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_uint64, c_uint64,
                                     c_uint64, c_double, c_int64, c_int, POINTER(c_uint64), _c_int_p]),
    'qfb_plan_refine_tile_lookahead': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                               c_uint64, c_uint64, c_uint64, c_double, c_int64, c_int, c_int,
                                               POINTER(c_uint64), _c_int_p]),
    'qfb_plan_set_threads': (c_int, [c_int]),
    'qfb_plan_split_rounds': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                      c_double, c_int, c_void_p, c_void_p, c_int, _c_int_p, POINTER(c_double),
                                      _c_int_p]),
    'qfb_batch_rho1': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'qfb_batch_rho1_workspace': (c_size_t, [c_int, c_int]),
    'qfb_batch_apply1': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'qfb_remap_swap': (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_uint64), c_void_p]),
    'qfb_remap_swap_slice': (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_uint64), c_int,
                                     POINTER(c_int), c_uint64, c_int, c_void_p]),
    'qfb_peer_barrier': (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, ctypes.c_uint32, c_void_p, c_void_p]),
    'qfb_gate_grad': (c_int, [c_void_p, c_void_p, c_int, c_int, _c_int_p, c_void_p, c_void_p]),
}


class QfbError(RuntimeError):
    """A libqfb200 call returned a non-zero status."""


_LIB = None


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libqfb200.so (building it in-tree first when it is absent or stale and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if build_if_missing and not _build.library_is_current():
        try:
            _build.build_library()
        except Exception as exc:  # stale-but-present library is still usable; absent is fatal
            if not os.path.exists(path):
                raise ImportError('libqfb200.so is missing and could not be built: {}'.format(exc)) from exc
    if not os.path.exists(path):
        raise ImportError('libqfb200.so not found at {} (run `python -m quantumflow_b200._build`)'.format(path))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so and the header disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().qfb_last_error()
        raise QfbError('libqfb200 error {}: {}'.format(status, msg.decode() if msg else '?'))


def int_array(values):
    vals = [int(v) for v in values]
    return (c_int * max(1, len(vals)))(*vals) if vals else (c_int * 1)(0)

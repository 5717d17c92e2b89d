"""The C-ABI library loads and exports every symbol that include/qfb200.h declares (no compute calls)."""
import os
import re

from quantumflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'qfb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qfb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_bound_functions():
    declared = _declared_symbols()
    assert declared, 'no declarations parsed'
    assert sorted(_lib.PROTOTYPES) == declared


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_string_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.qfb_version() >= 100
    assert isinstance(lib.qfb_last_error(), bytes)


def test_bad_arguments_are_reported_not_executed():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    import ctypes
    lib = _lib.load()
    bits = _lib.int_array([0])
    mat = (ctypes.c_double * 8)()
    rc = lib.qfb_apply_dense(None, None, 3, mat, 1, bits, 0, bits, 0, None)
    assert rc == 1 and b'null' in lib.qfb_last_error()
    dummy = ctypes.c_void_p(16)
    rc = lib.qfb_apply_dense(dummy, dummy, 3, mat, 1, _lib.int_array([5]), 0, bits, 0, None)
    assert rc == 1 and b'not local' in lib.qfb_last_error()
    rc = lib.qfb_plan_upload(ctypes.c_char_p(b'x' * 64), 64, ctypes.byref(ctypes.c_void_p()), None)
    assert rc == 1 and b'plan' in lib.qfb_last_error()

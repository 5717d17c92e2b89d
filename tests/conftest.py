"""Test configuration.

`-m "not gpu"`: oracle vs golden fixtures, host-side logic (gate algebra, planner + plan emulator), C-ABI load /
export checks -- runs anywhere in a few minutes.
`-m gpu`: parity tests proper; every one of them calls the CUDA kernels through the C ABI on cuda:0 and compares
with the oracle / the reference-generated fixtures. They never read /root/reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)

GOLDEN = os.path.join(TESTS, 'golden')

# amplitude parity bar of BASELINE.json's north_star (complex128, max-abs)
AMP_TOL = 1e-10


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); runs the kernels through the C ABI')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:       # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope='session', autouse=True)
def _built_libraries():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the C oracle once per session."""
    from quantumflow_b200 import _build
    try:
        _build.build_library()
    except Exception as exc:            # pragma: no cover
        if not os.path.exists(_build.LIB_PATH):
            pytest.exit('libqfb200.so missing and not buildable: {}'.format(exc), returncode=3)
    from oracle import c_oracle
    c_oracle.build()
    yield

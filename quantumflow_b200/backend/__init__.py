"""Backend selection (reference: quantumflow/backend/__init__.py:10-39). This package ships exactly one backend,
`b200`; the module-level names are the reference's backend contract so `from quantumflow_b200 import backend
as bk` is a drop-in for `from quantumflow import backend as bk`."""
from ..config import BACKEND, SEED          # noqa: F401
from .b200bk import *                       # noqa: F401,F403
from .b200bk import __all__ as _bk_all
from .b200bk import set_random_seed as _set_random_seed

__all__ = list(_bk_all) + ['BACKEND', 'SEED']

if SEED is not None:
    _set_random_seed(SEED)

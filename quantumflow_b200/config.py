"""Environment-driven configuration (same variable names as the reference, quantumflow/config.py:19-62).

QUANTUMFLOW_BACKEND   only 'b200' is accepted here (the reference's value set is extended by one)
QUANTUMFLOW_SEED      seeds `random`, numpy's global RandomState (the shared RNG stream) and torch
QUANTUMFLOW_LOG       log level of the 'quantumflow' logger
QUANTUMFLOW_DEVICE    CUDA device index for amplitude tensors (default: torch's current device)
"""
import logging
import os
import random

_PREFIX = 'QUANTUMFLOW_'

version = '0.1.0'

logging.getLogger('quantumflow').addHandler(logging.StreamHandler())
_LOGLEVEL = os.getenv(_PREFIX + 'LOG', None)
if _LOGLEVEL is not None:
    logging.getLogger('quantumflow').setLevel(_LOGLEVEL)

DEFAULT_BACKEND = 'b200'
BACKENDS = ('b200',)
BACKEND = os.getenv(_PREFIX + 'BACKEND', DEFAULT_BACKEND)
if BACKEND not in BACKENDS:
    raise ValueError('Unknown backend: {}BACKEND={} (this package only provides "b200")'.format(_PREFIX, BACKEND))
logging.getLogger('quantumflow').info('QuantumFlow Backend: %s', BACKEND)

TOLERANCE = 1e-6
"""Tolerance used in floating point comparisons (reference value)."""

_ENVSEED = os.getenv(_PREFIX + 'SEED', None)
SEED = int(_ENVSEED) if _ENVSEED is not None else None
if SEED is not None:
    random.seed(SEED)

_ENVDEV = os.getenv(_PREFIX + 'DEVICE', None)
DEVICE_INDEX = int(_ENVDEV) if (_ENVDEV is not None and _ENVDEV.isdigit()) else None

"""Load the reference's hot-path modules from /root/reference WITHOUT its package __init__ (which imports
cvxpy / pyquil, both absent). Only usable in the build container: the reference tree does not exist on the
GPU box, so nothing under `-m gpu`, smoke() or bench.py may import this. Used by make_golden.py only.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('QF_REFERENCE_ROOT', '/root/reference')

_MODULES = ['config', 'backend', 'cbits', 'qubits', 'states', 'utils', 'ops', 'stdops', 'gates', 'stdgates',
            'channels', 'circuits', 'paulialgebra', 'programs', 'dagcircuit', 'qaoa']


def load_reference():
    """Return a module object that behaves like `import quantumflow as qf` for the hot path."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'quantumflow')):
        raise ImportError('reference tree not found at ' + REFERENCE_ROOT)
    if 'quantumflow' in sys.modules and getattr(sys.modules['quantumflow'], '_qf_shim', False):
        return sys.modules['quantumflow']
    if os.environ.get('QUANTUMFLOW_BACKEND', 'numpy') not in ('numpy', 'b200'):
        raise ImportError('refshim needs the reference numpy backend')
    saved = os.environ.pop('QUANTUMFLOW_BACKEND', None)   # reference default = numpy
    pkg = types.ModuleType('quantumflow')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'quantumflow')]
    pkg.__version__ = '?.?.?'
    pkg._qf_shim = True
    sys.modules['quantumflow'] = pkg
    sys.dont_write_bytecode = True   # the reference mount is read-only
    for name in _MODULES:
        mod = importlib.import_module('quantumflow.' + name)
        setattr(pkg, name, mod)
        exported = getattr(mod, '__all__', None)
        if exported is None:      # qaoa.py has no __all__
            exported = [s for s in vars(mod) if not s.startswith('_')] if name == 'qaoa' else []
        for sym in exported:
            if hasattr(mod, sym) and not sym.startswith('__'):
                setattr(pkg, sym, getattr(mod, sym))
    if saved is not None:
        os.environ['QUANTUMFLOW_BACKEND'] = saved
    return pkg

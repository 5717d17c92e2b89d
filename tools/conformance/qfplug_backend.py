"""pytest plugin (-p qfplug_backend): the REFERENCE's own classes on the b200 backend.

Loads the reference's hot-path modules from QF_REFERENCE_ROOT exactly like the SURVEY Appendix-C shim, but with
`quantumflow.backend` = quantumflow_b200.backend.b200ref -- what the 3-line patch of INTEGRATION.md section 2
does (`QUANTUMFLOW_BACKEND=b200`). The reference's unmodified test files then exercise qf.State / Gate.run /
Circuit.run / Channel.evolve of the reference itself with every amplitude tensor in HBM and every tensormul a
libqfb200 kernel; the session summary prints the library's launch counter (0 would mean a CPU path).
Test infrastructure: nothing in the product imports this."""
import importlib
import os
import sys
import types

ROOT = os.environ.get('QF_REFERENCE_ROOT', '/root/reference')
_MODULES = ['cbits', 'qubits', 'states', 'utils', 'ops', 'stdops', 'gates', 'stdgates', 'channels', 'circuits',
            'paulialgebra', 'programs', 'measures', 'decompositions', 'dagcircuit', 'qaoa']


def _load():
    if getattr(sys.modules.get('quantumflow'), '_qf_shim', None) == 'backend':
        return
    sys.dont_write_bytecode = True
    sys.modules.setdefault('cvxpy', types.ModuleType('cvxpy'))     # measures.py imports it for the diamond norm only
    try:
        import PIL.Image                                           # noqa: F401  (visualization attribute at def time)
    except ImportError:
        pass
    os.environ.pop('QUANTUMFLOW_BACKEND', None)          # the reference's config.py rejects values it does not know
    pkg = types.ModuleType('quantumflow')
    pkg.__path__ = [os.path.join(ROOT, 'quantumflow')]
    pkg.__version__ = '?.?.?'
    pkg._qf_shim = 'backend'
    sys.modules['quantumflow'] = pkg
    cfg = importlib.import_module('quantumflow.config')
    cfg.BACKEND = 'b200'                                 # INTEGRATION.md section 2, first hunk of the patch
    pkg.config = cfg
    # second hunk: `from quantumflow_b200.backend.b200ref import *` inside the reference's backend package (the
    # package object keeps the reference's path: qubits.py:18 imports EINSUM_SUBSCRIPTS from backend.numpybk)
    from quantumflow_b200.backend import b200ref
    bk = types.ModuleType('quantumflow.backend')
    bk.__path__ = [os.path.join(ROOT, 'quantumflow', 'backend')]
    for name in b200ref.__all__:
        setattr(bk, name, getattr(b200ref, name))
    bk.BACKEND, bk.SEED = 'b200', cfg.SEED
    sys.modules['quantumflow.backend'] = bk
    pkg.backend = bk
    for name in _MODULES:
        mod = importlib.import_module('quantumflow.' + name)
        setattr(pkg, name, mod)
        exported = getattr(mod, '__all__', None)
        if exported is None:
            exported = [s for s in vars(mod) if not s.startswith('_')] if name == 'qaoa' else []
        for sym in exported:
            if hasattr(mod, sym) and not sym.startswith('__'):
                setattr(pkg, sym, getattr(mod, sym))


_load()


def pytest_terminal_summary(terminalreporter):
    from quantumflow_b200 import engine
    terminalreporter.write_line('qfplug_backend: reference classes on b200ref; libqfb200 kernel launches = {}'.format(
        engine.launch_count()))

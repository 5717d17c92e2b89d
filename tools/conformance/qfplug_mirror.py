"""pytest plugin (-p qfplug_mirror): the mirror package under the reference's name.

`import quantumflow as qf`, `from quantumflow import backend as bk`, `from quantumflow.utils import ...` resolve to
quantumflow_b200, so the reference's unmodified hot-path test files run against the engine (planner, sweep
kernels, read-out kernels). The session summary prints the library's launch counter.
Test infrastructure: nothing in the product imports this."""
import importlib
import sys

import quantumflow_b200 as _qfb

sys.modules['quantumflow'] = _qfb
for _name in ('backend', 'config', 'cbits', 'qubits', 'states', 'utils', 'ops', 'stdops', 'gates', 'stdgates',
              'channels', 'circuits', 'programs', 'dagcircuit', 'qaoa', 'measures'):
    try:
        sys.modules['quantumflow.' + _name] = importlib.import_module('quantumflow_b200.' + _name)
    except ImportError:
        pass


def pytest_terminal_summary(terminalreporter):
    from quantumflow_b200 import engine
    terminalreporter.write_line('qfplug_mirror: quantumflow_b200 as quantumflow; libqfb200 kernel launches = {}'.format(
        engine.launch_count()))

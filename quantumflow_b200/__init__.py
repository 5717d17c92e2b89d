"""quantumflow_b200 -- a B200-native gate-application engine behind the QuantumFlow API.

`import quantumflow_b200 as qf` gives the hot-path subset of the reference's flat namespace
(quantumflow/__init__.py:5-23): states, gates, channels, circuits, programs, QAOA helpers and the closeness predicates,
with `qf.backend` being the b200 tensor backend. Optional-dependency modules of the reference (visualization,
datasets, cvxpy-based measures, the pyQuil half of forest) are out of scope and are not imported; `qf.forest` holds
the part of the reference's forest module that sits on the path: the QAM-shaped front of `Program.run`.
"""
from . import backend                       # noqa: F401
from .config import *                       # noqa: F401,F403
from .cbits import *                        # noqa: F401,F403
from .qubits import *                       # noqa: F401,F403
from .states import *                       # noqa: F401,F403
from .ops import *                          # noqa: F401,F403
from .gates import *                        # noqa: F401,F403
from .stdgates import *                     # noqa: F401,F403
from .channels import *                     # noqa: F401,F403
from .stdops import *                       # noqa: F401,F403
from .circuits import *                     # noqa: F401,F403
from .dagcircuit import *                   # noqa: F401,F403
from .programs import *                     # noqa: F401,F403
from .measures import *                     # noqa: F401,F403
from .qaoa import *                         # noqa: F401,F403
from .trajectories import StateBatch       # noqa: F401  (engine-only: batched stochastic trajectories)
from . import utils, workloads, planner, engine, classify, stateio, trajectories, forest   # noqa: F401

from .config import version as __version__  # noqa: F401


def __getattr__(name: str):
    # qf.Parameter is sympy's Symbol in the reference; sympy is only imported when somebody asks for it
    if name == 'Parameter':
        from .programs import __getattr__ as _lazy
        return _lazy(name)
    raise AttributeError("module 'quantumflow_b200' has no attribute {!r}".format(name))

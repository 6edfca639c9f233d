"""idocp_b200 -- B200-native batched optimal-control engine for idocp's Newton-step hot path.

Everything numerical runs in the CUDA library ``libidocp_b200.so`` (sm_100a) behind the C-ABI of
``include/idocp_b200.h``; this package only marshals arguments.  No CPU fallback exists.
"""
from .capi import Idocp_b200Error, Library, Problem, default_library  # noqa: F401
from .solvers import (DerivativeChecker, ShardedSolver, UnOCPSolver, UnParNMPCSolver, benchmark_problem, config_space_problem,  # noqa: F401
                      task_space_3d_problem, task_space_circle_ref, task_space_problem)

__version__ = "0.1"
from .hybrid import ContactSequence, OCPDiscretizer  # noqa: F401,E402
from .ocp_solver import OCPSolver  # noqa: F401,E402
from .capi import FbProblem  # noqa: F401,E402
from .mpc import BatchedMPC  # noqa: F401,E402

"""TaskSpace3DCost / TimeVaryingTaskSpace3DCost (src/cost/task_space_3d_cost.cpp,
time_varying_task_space_3d_cost.cpp) through both solvers: the library under test vs the oracle, bit for bit."""
import numpy as np

import idocp_b200 as I
from helpers import check_iteration, check_solution, make_pair


def moving_target(t):
    """A user's compute_q_3d_ref(t) in the 12-double table row format [R (unused), p]."""
    return np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0, 0.546, 0.1 * np.sin(np.pi * t), 0.76 + 0.05 * np.cos(np.pi * t)])


def fixed_target(t):
    return moving_target(0.0)


def run_task3d(lib, oracle, batch, iters, N=12, T=0.6, kinds=("unocp", "unparnmpc"), line_search=(False, True)):
    prob = I.task_space_3d_problem(lib, N=N, T=T)
    rng = np.random.default_rng(3)
    q0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]) + rng.uniform(-0.3, 0.3, (batch, 7))
    v0 = rng.uniform(-0.2, 0.2, (batch, 7))
    for kind in kinds:
        for ref in (moving_target, fixed_target):
            for ls in (line_search if kind == "unocp" else (False,)):
                solver, oracles = make_pair(I, oracle, lib, prob, q0, v0, kind=kind, task_ref=ref)
                first = last = None
                for it in range(iters):
                    k = check_iteration(solver, oracles, q0, v0, line_search=ls)
                    first = k if first is None else first
                    last = k
                check_solution(solver, oracles)
                if kind != "unocp":   # a failed factorisation (status bit 0) is reported exactly where the oracle reports one
                    assert np.array_equal((solver.getStatus() & 1) != 0, [o.chol_info() != 0 for o in oracles])
    return first, last

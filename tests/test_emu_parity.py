"""CUDA sources under the SIMT emulator vs the oracle (CPU-only container).

This exercises the exact kernel code (idocp_b200/csrc/*.cuh: octet shuffles, shared-memory tiles,
launch geometry, C-ABI staging) with a g++ build in which every CUDA thread is a fiber.  It is a
check of the lane-parallel algorithms, not the product path: the GPU parity proper is
tests/test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

import idocp_b200 as I
from conftest import make_states
from helpers import DIR_FIELDS, SOL_FIELDS, check_iteration, check_solution, make_pair, rel_close


def test_emulator_library_is_not_the_default():
    assert "emu" not in I.capi.DEFAULT_LIBRARY


@pytest.mark.parametrize("batch", [1, 5, 8])
def test_unocp_iterations_match_oracle(emu_lib, oracle, batch):
    prob = I.benchmark_problem(emu_lib)
    q0, v0 = make_states(batch, 100 + batch)
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0)
    for it in range(4):
        check_iteration(solver, oracles, q0, v0)
    check_solution(solver, oracles)
    for name in ("slack", "dual"):
        x = solver.getConstraintData(name)
        ref = np.array([o.get_constraint_data(name) for o in oracles])
        assert np.array_equal(x, ref), name
    assert np.all(solver.getStatus() == 0)
    assert np.array_equal(solver.isCurrentSolutionFeasible(), [o.is_feasible() for o in oracles])


def test_unocp_condensed_kkt_matches_oracle(emu_lib, oracle):
    """k_linearize output (21x21 Q, 35 residual) vs SplitUnOCP::linearizeOCP of the oracle.
    The oracle's Riccati sweep updates Q in place, so compare the last stage, whose update adds only
    the diagonal terminal P (dt^2-scaled), by undoing it, and Fx / lq / lv everywhere."""
    prob = I.benchmark_problem(emu_lib)
    q0, v0 = make_states(3, 7)
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0)
    solver.setPipelining(False)    # getUnKKT must show the linearisation the direction came from, like the oracle's
    check_iteration(solver, oracles, q0, v0)       # move off the initial guess
    solver.updateSolution(0.0, q0, v0)
    for b, o in enumerate(oracles):
        o.update_solution(0.0, q0[b], v0[b])
    dt = prob.T / prob.N
    for stage in (0, 1, 2, prob.N - 1):
        Q, res = solver.getUnKKT(stage)
        for b, o in enumerate(oracles):
            Qo, ro = o.get_unkkt(stage)
            assert np.array_equal(res[b][:14], ro[:14])     # Fq, Fv
            assert np.array_equal(res[b][21:], ro[21:])     # condensed lq, lv
            if stage == prob.N - 1:
                Pqq, Pvv = np.diag([10.0] * 7), np.diag([0.1] * 7)
                Qo = Qo.copy()
                Qo[7:14, 7:14] -= Pqq
                Qo[7:14, 14:21] -= dt * Pqq
                Qo[14:21, 14:21] -= dt * dt * Pqq + Pvv
                Qo[0:7, 14:21] -= dt * Pvv
                Qo[0:7, 0:7] -= dt * dt * Pvv
                Qo[14:21, 7:14] = 0.0     # Qvq is only formed by the Riccati step
                scale = np.max(np.abs(Qo))
                assert np.max(np.abs(Q[b] - Qo)) <= 1e-9 * scale


def test_filter_line_search_matches_oracle(emu_lib, oracle):
    """updateSolution(..., line_search=True): UnLineSearch + LineSearchFilter (unline_search.hpp:62-91)
    as lock-step batched rounds; step sizes (accepted / backtracked / floored) bit-identical."""
    import bench
    prob = I.benchmark_problem(emu_lib)
    q0, v0 = bench.initial_states(100, 5, list(prob.q_min), list(prob.q_max))
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0)
    seen = set()
    for it in range(6):
        check_iteration(solver, oracles, q0, v0, line_search=True)
        p, _ = solver.getStepSizes()
        assert np.all(p >= 0.05) and np.all(p <= 1.0)
        seen.update(np.round(p, 6).tolist())
    assert 0.05 in seen and len(seen) > 2        # floor reached and non-trivial steps taken
    check_solution(solver, oracles)
    # clearLineSearchFilter on both sides keeps them in lock-step
    solver.clearLineSearchFilter()
    for o in oracles:
        o.clear_line_search_filter()
    check_iteration(solver, oracles, q0, v0, line_search=True)


def test_config_space_problem_long_horizon(emu_lib, oracle):
    """BASELINE configs[0] shape (N=60, T=3) for a couple of iterations."""
    prob = I.config_space_problem(emu_lib)
    q0 = np.array([[np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2]])
    v0 = np.zeros((1, 7))
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0)
    for _ in range(2):
        check_iteration(solver, oracles, q0, v0)
    check_solution(solver, oracles)


@pytest.mark.parametrize("batch,N", [(1, 1), (2, 2), (5, 6)])
def test_unparnmpc_iterations_match_oracle(emu_lib, oracle, batch, N):
    """UnParNMPCSolver (backward-Euler stages, explicit 35x35 KKT inverse per stage, backward/forward
    correction sweeps): directions, step sizes, KKT error and iterate bit-identical to the oracle."""
    prob = I.benchmark_problem(emu_lib, N=N, T=0.05 * N)
    q0, v0 = make_states(batch, 300 + batch)
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0, kind="unparnmpc")
    for it in range(3):
        check_iteration(solver, oracles, q0, v0)
    check_solution(solver, oracles)
    assert np.all(solver.getStatus() == 0)


@pytest.mark.parametrize("kind,N,line_search", [("unocp", 4, False), ("unparnmpc", 4, False), ("unocp", 3, True)])
def test_task_space_cost_matches_oracle(emu_lib, oracle, kind, N, line_search):
    """TimeVaryingTaskSpace6DCost (examples/iiwa14/task_space_ocp.cpp problem, short horizon): frame
    kinematics, log6 / Jlog6, dense Gauss-Newton Hessian (also at the terminal stage and in the ParNMPC
    aux matrix), line-search cost -- bit-identical to the oracle."""
    prob = I.task_space_problem(emu_lib, N=N, T=0.05 * N)
    rng = np.random.default_rng(3)
    q0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]) + rng.uniform(-0.3, 0.3, (3, 7))
    v0 = rng.uniform(-0.2, 0.2, (3, 7))
    solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0, kind=kind, task_ref=I.task_space_circle_ref)
    for it in range(3):
        check_iteration(solver, oracles, q0, v0, line_search=line_search)
    check_solution(solver, oracles)


@pytest.mark.parametrize("task", [False, True])
def test_unparnmpc_filter_line_search_matches_oracle(emu_lib, oracle, task):
    """UnLineSearch::computeStepSize<UnParNMPC> (unline_search.hpp:62-91, unline_search.cpp:87-122): backward-Euler
    defects against the TRIAL previous stage, terminal cost on the last stage, last-stage reference time t + N dt."""
    if task:
        prob = I.task_space_problem(emu_lib, N=4, T=0.2)
        rng = np.random.default_rng(3)
        q0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]) + rng.uniform(-0.3, 0.3, (3, 7))
        v0 = rng.uniform(-0.2, 0.2, (3, 7))
        solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0, kind="unparnmpc", task_ref=I.task_space_circle_ref)
    else:
        prob = I.config_space_problem(emu_lib)
        prob.N, prob.T = 5, 0.25
        rng = np.random.default_rng(4)
        q0, v0 = rng.uniform(-1, 1, (3, 7)), rng.uniform(-0.2, 0.2, (3, 7))
        solver, oracles = make_pair(I, oracle, emu_lib, prob, q0, v0, kind="unparnmpc")
    steps = []
    for it in range(4):
        check_iteration(solver, oracles, q0, v0, line_search=True)
        steps.append(solver.getStepSizes()[0])
    steps = np.array(steps)
    assert steps.min() == 0.05 and len(np.unique(steps)) > 3
    check_solution(solver, oracles)
    solver.clearLineSearchFilter()
    for o in oracles:
        o.clear_line_search_filter()
    check_iteration(solver, oracles, q0, v0, line_search=True)


def with_acceleration_limits(prob, a_limit):
    """JointAccelerationLowerLimit(robot, -a_limit) + JointAccelerationUpperLimit(robot, +a_limit)."""
    prob.enable_acceleration_limit[0] = prob.enable_acceleration_limit[1] = 1
    for j in range(7):
        prob.a_min[j], prob.a_max[j] = -a_limit, a_limit
    return prob


def acceleration_limit_scenario(lib, oracle, kind, task, line_search, batch=3, iters=4):
    """The two acceleration-limit components through the ACC kernel instantiations vs the oracle: directions, step sizes,
    KKT errors, iterates and the slack / dual rows of all eight components, bit for bit."""
    rng = np.random.default_rng(21)
    if task:
        prob = with_acceleration_limits(I.task_space_problem(lib, N=4, T=0.2), 8.0)
        q0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]) + rng.uniform(-0.3, 0.3, (batch, 7))
        v0 = rng.uniform(-0.2, 0.2, (batch, 7))
        solver, oracles = make_pair(I, oracle, lib, prob, q0, v0, kind=kind, task_ref=I.task_space_circle_ref)
    else:
        prob = with_acceleration_limits(I.benchmark_problem(lib, N=6, T=0.3), 25.0)
        q0, v0 = make_states(batch, 77)
        solver, oracles = make_pair(I, oracle, lib, prob, q0, v0, kind=kind)

    def rows():
        if kind != "unocp":    # the oracle's UnParNMPCSolver has no constraint-data getter; its rows drive every compared quantity
            return
        for name in ("slack", "dual", "acc_slack", "acc_dual"):
            x = solver.getConstraintData(name)
            ref = np.array([o.get_constraint_data(name) for o in oracles])
            assert np.array_equal(x, ref), name
    rows()
    assert np.all(solver.getConstraintData("acc_slack") > 0)
    for it in range(iters):
        check_iteration(solver, oracles, q0, v0, line_search=line_search)
        rows()
    check_solution(solver, oracles)
    if kind == "unocp":
        assert np.array_equal(solver.isCurrentSolutionFeasible(), [o.is_feasible() for o in oracles])
    return solver, oracles


@pytest.mark.parametrize("kind,task,line_search", [("unocp", False, False), ("unocp", False, True), ("unparnmpc", False, False),
                                                   ("unocp", True, False), ("unparnmpc", True, True)])
def test_acceleration_limits_match_oracle(emu_lib, oracle, kind, task, line_search):
    acceleration_limit_scenario(emu_lib, oracle, kind, task, line_search)


def test_acceleration_limits_force_the_literal_sequence(emu_lib):
    prob = with_acceleration_limits(I.benchmark_problem(emu_lib, N=3, T=0.15), 25.0)
    s = I.UnOCPSolver(prob, 2, lib=emu_lib)
    s.setPipelining(True)            # ignored: the fused update + linearisation carries the six joint-limit components only
    q0, v0 = make_states(2, 5)
    s.setSolution("q", q0)
    s.setSolution("v", v0)
    s.updateSolution(0.0, q0, v0)
    prof = s.launchCount()
    assert prof > 0
    plain = I.UnOCPSolver(I.benchmark_problem(emu_lib, N=3, T=0.15), 2, lib=emu_lib)
    with pytest.raises(I.Idocp_b200Error, match="not enabled"):
        plain.getConstraintData("acc_slack")


def test_error_paths(emu_lib):
    prob = I.benchmark_problem(emu_lib)
    bad = I.benchmark_problem(emu_lib)
    bad.N = 0
    with pytest.raises(I.Idocp_b200Error, match="N must be positive"):
        I.UnOCPSolver(bad, 2, lib=emu_lib)
    bad = I.benchmark_problem(emu_lib)
    bad.T = -1.0
    with pytest.raises(I.Idocp_b200Error, match="T must be positive"):
        I.UnOCPSolver(bad, 2, lib=emu_lib)
    s = I.UnOCPSolver(prob, 2, lib=emu_lib)
    with pytest.raises(I.Idocp_b200Error, match="name must be q, v, a, or u"):
        s.setSolution("lmd", np.zeros(7))
    with pytest.raises(I.Idocp_b200Error):
        s.getSolution("nope")
    with pytest.raises(ValueError):
        s.updateSolution(0.0, np.zeros((3, 7)), np.zeros((3, 7)))


def test_pipelined_update_equals_literal_sequence(emu_lib, oracle):
    """k_step_min + k_update_linearize (persistent: update + linearisation of the new iterate kept for the next call) vs the literal
    linearise / Riccati / expand / update sequence: identical bits, also across setSolution (which invalidates the kept
    linearisation), computeKKTResidual between the calls, a changed x0, and the filter line search."""
    prob = I.benchmark_problem(emu_lib)
    q0, v0 = make_states(5, 11)
    a = I.UnOCPSolver(prob, 5, lib=emu_lib)
    b = I.UnOCPSolver(prob, 5, lib=emu_lib)
    b.setPipelining(False)
    for s in (a, b):
        s.setSolution("q", q0)
        s.setSolution("v", v0)

    def same():
        for name in SOL_FIELDS:
            assert np.array_equal(a.getSolution(name), b.getSolution(name), equal_nan=True), name
        for name in DIR_FIELDS:
            assert np.array_equal(a.getDirection(name), b.getDirection(name), equal_nan=True), name
        assert np.array_equal(a.getConstraintData("slack"), b.getConstraintData("slack"), equal_nan=True)
        assert np.array_equal(a.getConstraintData("dual"), b.getConstraintData("dual"), equal_nan=True)
        assert np.array_equal(a.getStepSizes(), b.getStepSizes(), equal_nan=True)
    for it in range(3):
        for s in (a, b):
            s.updateSolution(0.0, q0, v0)
        same()
    for s in (a, b):                       # KKT residual in between does not disturb the kept linearisation
        s.computeKKTResidual(0.0, q0, v0)
    assert np.array_equal(a.KKTError(), b.KKTError())
    q1, v1 = q0 + 0.01, v0 - 0.02          # MPC: the measured state moves, the linearisation is still valid
    for s in (a, b):
        s.updateSolution(0.0, q1, v1)
    same()
    for s in (a, b):                       # setSolution re-initialises the constraints: the kept linearisation is stale
        s.setSolution("q", q1)
        s.updateSolution(0.0, q1, v1)
    same()
    for it in range(2):
        for s in (a, b):
            s.updateSolution(0.0, q1, v1, True)
        same()
    assert a.launchCount() > 0 and b.launchCount() > 0


def kkt_by_product_scenario(lib, task):
    """computeKKTResidual after a pipelined updateSolution: from the second call on the fused update + linearisation leaves the
    squared stage norms of the new iterate behind (k_update_linearize<.., KKT>) and computeKKTResidual only sums them; the
    errors equal the literal sequence's bit for bit -- across setSolution, a moved x0 and the line search."""
    if task:
        prob = I.task_space_problem(lib, N=4, T=0.2)
        rng = np.random.default_rng(5)
        q0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]) + rng.uniform(-0.3, 0.3, (5, 7))
        v0 = rng.uniform(-0.2, 0.2, (5, 7))
    else:
        prob = I.benchmark_problem(lib)
        q0, v0 = make_states(5, 19)
    a = I.UnOCPSolver(prob, 5, lib=lib)
    b = I.UnOCPSolver(prob, 5, lib=lib)
    b.setPipelining(False)
    for s in (a, b):
        s.setSolution("q", q0)
        s.setSolution("v", v0)
        if task:
            s.setTaskReference(I.task_space_circle_ref, 0.0)
    launches = []
    for it in range(5):
        for s in (a, b):
            s.updateSolution(0.0, q0, v0, it == 3)
        n0 = a.launchCount()
        for s in (a, b):
            s.computeKKTResidual(0.0, q0, v0)
        launches.append(a.launchCount() - n0)
        assert np.array_equal(a.KKTError(), b.KKTError(), equal_nan=True), it
        for name in SOL_FIELDS:
            assert np.array_equal(a.getSolution(name), b.getSolution(name), equal_nan=True), (it, name)
    assert launches[0] == 2 and launches[1:] == [1, 1, 1, 1], launches      # k_linearize<residual> + k_kkt_sum, then k_kkt_sum alone
    for s in (a, b):                       # setSolution invalidates the kept linearisation and its by-product
        s.setSolution("q", q0 + 0.01)
        s.computeKKTResidual(0.0, q0, v0)
    assert np.array_equal(a.KKTError(), b.KKTError(), equal_nan=True)
    for s in (a, b):
        s.updateSolution(0.0, q0 + 0.02, v0)
        s.computeKKTResidual(0.0, q0 + 0.02, v0)
    assert np.array_equal(a.KKTError(), b.KKTError(), equal_nan=True)


@pytest.mark.parametrize("task", [False, True])
def test_kkt_error_as_by_product_of_the_pipelined_update(emu_lib, task):
    kkt_by_product_scenario(emu_lib, task)

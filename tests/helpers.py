"""Shared parity-check logic: drive the batched engine (emulator or GPU) and the oracle side by side."""
import numpy as np

RTOL = 1e-9        # north_star: Newton direction / trajectory within 1e-9 relative
KKT_ATOL = 1e-8    # north_star: 1e-8 absolute on the KKT error
SOL_FIELDS = ["q", "v", "a", "u", "beta", "lmd", "gmm"]
DIR_FIELDS = ["dq", "dv", "da", "du", "dbeta", "dlmd", "dgmm"]


def rel_close(x, ref, rtol=RTOL, floor=1e-12):
    """max |x - ref| <= rtol * max|ref| (per field, scaled by the field's magnitude) ."""
    scale = max(float(np.max(np.abs(ref))), floor)
    return float(np.max(np.abs(x - ref))) <= rtol * scale + floor


def copy_problem(src, dst):
    """Copy the fields shared by idocp_b200.Problem and oracle_py.Problem."""
    for name in ("N", "T", "barrier", "fraction_rate", "task_enabled"):
        setattr(dst, name, getattr(src, name))
    for name in ("task_q_weight", "task_qf_weight"):
        a, b = getattr(src, name), getattr(dst, name)
        for i in range(6):
            b[i] = a[i]
    for name in ("q_ref", "v_ref", "u_ref", "q_weight", "v_weight", "a_weight", "u_weight", "qf_weight",
                 "vf_weight", "q_min", "q_max", "v_max", "u_max"):
        a, b = getattr(src, name), getattr(dst, name)
        for i in range(7):
            b[i] = a[i]
    # JointAcceleration{Lower,Upper}Limit: enable flags and bounds (the two structures name the flags differently)
    for k in range(2):
        dst.enable_acc[k] = src.enable_acceleration_limit[k]
    for name in ("a_min", "a_max"):
        a, b = getattr(src, name), getattr(dst, name)
        for i in range(7):
            b[i] = a[i]
    return dst


def make_pair(I, O, lib, problem, q0, v0, kind="unocp", task_ref=None, t=0.0):
    """task_ref: the user's compute_q_6d_ref(t) -> 12 doubles, sampled by both sides at the stage times."""
    batch = q0.shape[0]
    cls = I.UnOCPSolver if kind == "unocp" else I.UnParNMPCSolver
    solver = cls(problem, batch, lib=lib)
    solver.setSolution("q", q0)
    solver.setSolution("v", v0)
    oprob = copy_problem(problem, O.default_problem())
    ocls = O.UnOCPSolver if kind == "unocp" else O.UnParNMPCSolver
    oracles = [ocls(oprob) for _ in range(batch)]
    for b, o in enumerate(oracles):
        o.set_solution("q", q0[b])
        o.set_solution("v", v0[b])
    if task_ref is not None:
        solver.setTaskReference(task_ref, t)
        table = O.task_ref_table(task_ref, t, problem.T, problem.N, kind)
        for o in oracles:
            o.set_task_ref(table)
    if kind != "unocp":   # examples/iiwa14/unparnmpc_benchmark.cpp:49-51
        solver.initBackwardCorrection(0.0)
        for o in oracles:
            o.init_backward_correction(0.0)
    return solver, oracles


def _same(x, ref, exact):
    """exact: bit-identical (the kernels and the oracle share one canonical IEEE-754 operation
    sequence, see idocp_b200/csrc/octet.cuh); otherwise the north_star tolerance."""
    return np.array_equal(x, ref, equal_nan=True) if exact else rel_close(x, ref)


def check_iteration(solver, oracles, q0, v0, t=0.0, line_search=False, check_direction=True, exact=True):
    """One updateSolution + computeKKTResidual on both sides; asserts parity; returns the KKT errors."""
    solver.updateSolution(t, q0, v0, line_search)
    for b, o in enumerate(oracles):
        o.update_solution(t, q0[b], v0[b], line_search)
    if check_direction:
        for name in DIR_FIELDS:
            d = solver.getDirection(name)
            ref = np.array([o.get_direction(name) for o in oracles])
            assert _same(d, ref, exact), "direction %s: max diff %g (scale %g)" % (
                name, np.nanmax(np.abs(d - ref)), np.nanmax(np.abs(ref)))
        p, dd = solver.getStepSizes()
        ref = np.array([o.step_sizes() for o in oracles])
        if exact:
            assert np.array_equal(p, ref[:, 0], equal_nan=True), (p, ref[:, 0])
            assert np.array_equal(dd, ref[:, 1], equal_nan=True), (dd, ref[:, 1])
        else:
            assert np.allclose(p, ref[:, 0], rtol=RTOL, atol=0), (p, ref[:, 0])
            assert np.allclose(dd, ref[:, 1], rtol=RTOL, atol=0), (dd, ref[:, 1])
    solver.computeKKTResidual(t, q0, v0)
    kkt = solver.KKTError()
    ref = []
    for b, o in enumerate(oracles):
        o.compute_kkt_residual(t, q0[b], v0[b])
        ref.append(o.kkt_error())
    ref = np.array(ref)
    if exact:
        assert np.array_equal(kkt, ref, equal_nan=True), (kkt, ref, kkt - ref)
    else:
        assert np.all(np.abs(kkt - ref) <= KKT_ATOL + RTOL * np.abs(ref)), (kkt, ref)
    return kkt, ref


def check_solution(solver, oracles, exact=True):
    for name in SOL_FIELDS:
        x = solver.getSolution(name)
        ref = np.array([o.get_solution(name) for o in oracles])
        assert _same(x, ref, exact), "solution %s: max diff %g (scale %g)" % (
            name, np.max(np.abs(x - ref)), np.max(np.abs(ref)))

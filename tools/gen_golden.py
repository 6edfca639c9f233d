#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/ FROM THE ORACLE.

The reference cannot be built in this image (no Eigen/Boost/pinocchio/urdfdom) and its own test
suite holds no known-answer vectors, so these fixtures pin the ORACLE (and,
through the parity tests, the CUDA kernels) against regressions; they are not outputs of upstream
idocp.  Re-run with:  python tools/gen_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import oracle_py as O  # noqa: E402


def run_unocp(problem, q0, v0, iters, keep_dir_iters=(0, 1)):
    s = O.UnOCPSolver(problem)
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    s.compute_kkt_residual(0.0, q0, v0)
    rec = {"q0": list(map(float, q0)), "v0": list(map(float, v0)), "kkt": [s.kkt_error()], "primal": [], "dual": [],
           "directions": {}}
    for it in range(iters):
        s.update_solution(0.0, q0, v0, False)
        st = s.step_sizes()
        rec["primal"].append(float(st[0]))
        rec["dual"].append(float(st[1]))
        if it in keep_dir_iters:
            rec["directions"][str(it)] = {n: s.get_direction(n).tolist() for n in ("dq", "dv", "da", "du", "dlmd", "dgmm", "dbeta")}
        s.compute_kkt_residual(0.0, q0, v0)
        rec["kkt"].append(s.kkt_error())
    rec["final"] = {n: s.get_solution(n).tolist() for n in ("q", "v", "a", "u", "lmd", "gmm", "beta")}
    return rec


def run_solver(kind, problem, q0, v0, iters, ref_fn=None, line_search=False):
    """KKT / step-size history and the final iterate of one oracle solve (UnOCPSolver or UnParNMPCSolver,
    optionally with the task-space reference table), the Convergence driver of ocp_benchmarker.hxx:37-51."""
    cls = O.UnOCPSolver if kind == "unocp" else O.UnParNMPCSolver
    s = cls(problem)
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    if ref_fn is not None:
        s.set_task_ref(O.task_ref_table(ref_fn, 0.0, problem.T, problem.N, kind))
    if kind != "unocp":
        s.init_backward_correction(0.0)
    s.compute_kkt_residual(0.0, q0, v0)
    rec = {"q0": list(map(float, q0)), "v0": list(map(float, v0)), "kkt": [s.kkt_error()], "primal": [], "dual": []}
    for it in range(iters):
        s.update_solution(0.0, q0, v0, line_search)
        st = s.step_sizes()
        rec["primal"].append(float(st[0]))
        rec["dual"].append(float(st[1]))
        if it == 0:
            rec["first_direction"] = {n: s.get_direction(n).tolist() for n in ("dq", "dlmd", "du")}
        s.compute_kkt_residual(0.0, q0, v0)
        rec["kkt"].append(s.kkt_error())
    rec["final"] = {n: s.get_solution(n).tolist() for n in ("q", "u", "lmd")}
    return rec


def main_solvers():
    """tests/golden/solvers_golden.json: UnParNMPCSolver and the task-space problem (BASELINE configs[1])."""
    out = {}
    # examples/iiwa14/unparnmpc_benchmark.cpp:22-56 single instance
    out["unparnmpc_benchmark_reference_instance"] = run_solver("unparnmpc", O.benchmark_problem(), np.full(7, 2.0),
                                                               np.zeros(7), 20)
    # config_space_ocp.cpp problem through the ParNMPC driver (converges)
    pc = O.config_space_problem()
    pc.N, pc.T = 20, 1.0
    qc = np.array([np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2])
    out["config_space_unparnmpc"] = run_solver("unparnmpc", pc, qc, np.zeros(7), 60)
    # examples/iiwa14/task_space_ocp.cpp:55-93 (T = 6, N = 120, q = (0, pi/2, ...), 30 iterations), UnOCPSolver as the
    # example and UnParNMPCSolver as BASELINE configs[1] asks
    pt = O.task_space_problem()
    qt = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0])
    out["task_space_ocp_unocp"] = run_solver("unocp", pt, qt, np.zeros(7), 30, ref_fn=O.task_space_ref)
    out["task_space_ocp_unparnmpc"] = run_solver("unparnmpc", pt, qt, np.zeros(7), 30, ref_fn=O.task_space_ref)
    pl = O.task_space_problem(N=20, T=1.0)
    out["task_space_line_search_unocp"] = run_solver("unocp", pl, qt + 0.2, np.zeros(7), 10, ref_fn=O.task_space_ref,
                                                     line_search=True)
    # task-space kinematics known answers
    rng = np.random.default_rng(2025)
    ts = []
    for _ in range(4):
        q = rng.uniform(-2.5, 2.5, 7)
        ref = O.task_space_ref(rng.uniform(0, 2))
        diff, JJ = O.task_evaluate(q, ref)
        R, p, J = O.frame_kinematics(q)
        ts.append({"q": q.tolist(), "ref": ref.tolist(), "diff": diff.tolist(), "JJ": JJ.tolist(), "R": R.tolist(),
                   "p": p.tolist(), "J": J.tolist()})
    out["task_space_kinematics"] = ts
    path = os.path.join(ROOT, "tests", "golden", "solvers_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    out = {}
    # BASELINE configs[2] single reference instance: examples/iiwa14/unocp_benchmark.cpp:44-52
    pb = O.benchmark_problem()
    out["unocp_benchmark_reference_instance"] = run_unocp(pb, np.full(7, 2.0), np.zeros(7), 50)
    # first instances of the bench batch (splitmix64 seed 20240001)
    q0, v0 = bench.initial_states(0, 4, list(pb.q_min), list(pb.q_max))
    out["unocp_benchmark_batch_head"] = [run_unocp(pb, q0[b], v0[b], 12) for b in range(4)]
    # BASELINE configs[0]: examples/iiwa14/config_space_ocp.cpp (T=3, N=60, 30 iterations)
    pc = O.config_space_problem()
    qc = np.array([np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2])
    out["config_space_ocp"] = run_unocp(pc, qc, np.zeros(7), 30, keep_dir_iters=(0,))
    # rigid-body known answers
    rng = np.random.default_rng(2024)
    rb = []
    for _ in range(4):
        q, v, a = rng.uniform(-2.5, 2.5, 7), rng.uniform(-4, 4, 7), rng.uniform(-8, 8, 7)
        dq, dv, da = O.rnea_derivatives(q, v, a)
        rb.append({"q": q.tolist(), "v": v.tolist(), "a": a.tolist(), "tau": O.rnea(q, v, a).tolist(),
                   "dtau_dq": dq.tolist(), "dtau_dv": dv.tolist(), "dtau_da": da.tolist()})
    out["rnea"] = rb
    path = os.path.join(ROOT, "tests", "golden", "unocp_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main_solvers()

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

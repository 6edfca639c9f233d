#!/usr/bin/env python3
"""Golden vectors of the ANYmal OCPSolver path, generated FROM THE ORACLE (the reference cannot be built here and
holds no known-answer vectors): examples/anymal/anymal_trotting.cpp, 25 iterations -- KKT history,
step sizes, the chain of stages and the final trajectory of a few stages.  Re-run: python tools/gen_golden_anymal.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import anymal_problems as ap  # noqa: E402
import fb_py  # noqa: E402
import oracle_py  # noqa: E402


def main():
    oracle_py.build()
    fb_py.lib()
    pr = ap.TrottingProblem()
    # the C++ example obtains the contact points from the product's host-side Robot::updateFrameKinematics
    # (idocp_b200_fb_contact_frame_positions, plain host arithmetic): use the very same numbers
    sys.path.insert(0, ROOT)
    import idocp_b200
    from idocp_b200 import capi
    lib = idocp_b200.default_library()
    pts = np.zeros((4, 3))
    lib.check(lib.L.idocp_b200_fb_contact_frame_positions(capi.dptr(np.ascontiguousarray(ap.Q_STANDING)), capi.dptr(pts)))
    assert np.allclose(pts, ap.standing_contact_points(fb_py), atol=1e-14)
    pr.standing_points = pts
    ocp = pr.make_oracle(fb_py)
    rec = {"chain": [[c["kind"], c["index"], c["t"], c["dt"], c["dimf"], c["dimi"]] for c in ocp.chain()], "kkt": [], "primal": [],
           "dual": [], "contact_points": pts.tolist()}
    ocp.compute_kkt_residual(0.0, pr.q0, pr.v0)
    rec["kkt"].append(ocp.kkt_error())
    for it in range(25):
        assert ocp.update_solution(0.0, pr.q0, pr.v0) == 0
        st = ocp.step_sizes()
        rec["primal"].append(float(st[0]))
        rec["dual"].append(float(st[1]))
        if it == 0:
            rec["first_direction"] = {str(e): {n: ocp.get(e, n).tolist() for n in ("dq", "dv", "du")} for e in (0, 11, 20)}
        ocp.compute_kkt_residual(0.0, pr.q0, pr.v0)
        rec["kkt"].append(ocp.kkt_error())
    n = len(ocp.chain())
    rec["final"] = {str(e): {nm: ocp.get(e, nm).tolist() for nm in ("q", "v", "u", "f")} for e in (0, 11, 20, n - 2)}
    rec["final"][str(n - 1)] = {nm: ocp.get(n - 1, nm).tolist() for nm in ("q", "v")}
    with open(os.path.join(ROOT, "tests", "golden", "anymal_trotting_golden.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("anymal_trotting: KKT %.3e -> %.3e" % (rec["kkt"][0], rec["kkt"][-1]))
    # examples/anymal/anymal_running.cpp, first 8 iterations (the example runs 350)
    rp = ap.RunningProblem(10)
    pts = np.zeros((4, 3))
    lib.check(lib.L.idocp_b200_fb_contact_frame_positions(capi.dptr(np.ascontiguousarray(rp.q_begin)), capi.dptr(pts)))
    rp.standing_points = pts
    ocp = rp.make_oracle(fb_py)
    ocp.set_threads(8)
    rec = {"chain": [[c["kind"], c["index"], c["t"], c["dt"], c["dimf"], c["dimi"]] for c in ocp.chain()], "kkt": [], "primal": [], "dual": []}
    ocp.compute_kkt_residual(0.0, rp.q0, rp.v0)
    rec["kkt"].append(ocp.kkt_error())
    for it in range(8):
        assert ocp.update_solution(0.0, rp.q0, rp.v0) == 0
        st = ocp.step_sizes()
        rec["primal"].append(float(st[0]))
        rec["dual"].append(float(st[1]))
        ocp.compute_kkt_residual(0.0, rp.q0, rp.v0)
        rec["kkt"].append(ocp.kkt_error())
    with open(os.path.join(ROOT, "tests", "golden", "anymal_running_golden.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("anymal_running: KKT %.3e -> %.3e, %d stages" % (rec["kkt"][0], rec["kkt"][-1], len(rec["chain"])))
    # examples/anymal/ocp_benchmark.cpp: standing, nonlinear FrictionCone, 10 iterations; and the same problem with
    # JointAcceleration{Lower,Upper}Limit (|a| <= 2 rad/s^2) pushed as well
    for tag, a_limit in (("", None), ("_acc", 2.0)):
        sp = ap.StandingBenchmarkProblem()
        if a_limit is not None:
            ap.with_nonlinear_cones_and_acceleration_limits(sp, cones=True, a_limit=a_limit)
            sp.problem.cone_nonlinear[1] = 0
        pts = np.zeros((4, 3))
        lib.check(lib.L.idocp_b200_fb_contact_frame_positions(capi.dptr(np.ascontiguousarray(ap.Q_STANDING)), capi.dptr(pts)))
        sp.standing_points = pts
        ocp = sp.make_oracle(fb_py)
        rec = {"kkt": [], "primal": [], "dual": []}
        ocp.compute_kkt_residual(0.0, sp.q0, sp.v0)
        rec["kkt"].append(ocp.kkt_error())
        for it in range(10):
            assert ocp.update_solution(0.0, sp.q0, sp.v0) == 0
            st = ocp.step_sizes()
            rec["primal"].append(float(st[0]))
            rec["dual"].append(float(st[1]))
            ocp.compute_kkt_residual(0.0, sp.q0, sp.v0)
            rec["kkt"].append(ocp.kkt_error())
        rec["final_f_stage0"] = ocp.get(0, "f").tolist()
        with open(os.path.join(ROOT, "tests", "golden", "anymal_ocp_benchmark%s_golden.json" % tag), "w") as f:
            json.dump(rec, f, indent=1)
        print("anymal_ocp_benchmark%s: KKT %.3e -> %.3e" % (tag, rec["kkt"][0], rec["kkt"][-1]), rec["primal"])


if __name__ == "__main__":
    main()

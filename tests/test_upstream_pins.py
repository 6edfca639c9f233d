"""Pins of the CPU oracle against REAL upstream idocp (mayataka/idocp + pinocchio + Eigen).

The reference cannot be built in this image (Eigen3 / Boost / pinocchio / urdfdom absent, no network) and ships no golden
vectors, so every parity claim of this repository is "CUDA == oracle" with the oracle validated only by independent
re-derivations (finite differences, Lagrangian dynamics, identities of the reference's unit tests).  tools/pin_against_idocp/
is the kit that closes the gap on a machine with idocp installed: `pin_dump` writes tests/golden/upstream_robot.json and
tests/golden/upstream_solvers.json, and THIS file checks the oracle against them with the north-star tolerances
(1e-9 relative, 1e-8 absolute on the KKT error).  Without the files the pins are SKIPPED -- loudly -- and the oracle stays
"parity unpinned" (DESIGN.md section 5).  The consumer itself is exercised on every run: the same checks are applied to a
file of the same schema produced from the oracle (a schema / plumbing test, not a pin)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

ROBOT_FILE = os.path.join(GOLDEN, "upstream_robot.json")
SOLVER_FILE = os.path.join(GOLDEN, "upstream_solvers.json")
MISSING = ("tests/golden/upstream_*.json not present: the oracle is NOT pinned to upstream idocp.  Build tools/pin_against_idocp "
           "against an idocp installation and run `pin_dump <idocp>/examples tests/golden` (see its CMakeLists.txt)")
RTOL, ATOL, KKT_ATOL = 1e-9, 1e-9, 1e-8


def close(x, y, rtol=RTOL, atol=ATOL):
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    return x.shape == y.shape and bool(np.all(np.abs(x - y) <= atol + rtol * np.maximum(np.abs(x), np.abs(y))))


# ---------------------------------------------------------------------------------------------------------------
# checkers (used on the upstream files and on the oracle-made self-test files)
# ---------------------------------------------------------------------------------------------------------------
def check_robot(rec, O, fb):
    bad = []
    for k, s in enumerate(rec["iiwa14"]):
        q, v, a = s["q"], s["v"], s["a"]
        dq, dv, da = O.rnea_derivatives(q, v, a)
        R, p, J = O.frame_kinematics(q)
        for name, got, ref in (("tau", O.rnea(q, v, a), s["tau"]), ("dtau_dq", dq, s["dtau_dq"]), ("dtau_dv", dv, s["dtau_dv"]),
                               ("dtau_da", da, s["dtau_da"]), ("frame22_position", p, s["frame22_position"]),
                               ("frame22_rotation", R, s["frame22_rotation"]), ("frame22_jacobian_local", J, s["frame22_jacobian_local"])):
            if not close(got, ref):
                bad.append(("iiwa14", k, name))
    for k, s in enumerate(rec["anymal"]):
        q, v, a = np.array(s["q"]), np.array(s["v"]), np.array(s["a"])
        active = [bool(x) for x in s["active"]]
        f = np.array(s["f"]).reshape(4, 3) * np.array(active, dtype=float)[:, None]
        tau, dq, dv, M = fb.rnea(q, v, a, f.ravel(), derivatives=True)
        for name, got, ref in (("tau", tau, s["tau"]), ("dtau_dq", dq, s["dtau_dq"]), ("dtau_dv", dv, s["dtau_dv"]),
                               ("dtau_da", M, s["dtau_da"])):
            if not close(got, ref):
                bad.append(("anymal", k, name))
        feet = [fb.contact(q, v, a, c, 0.05, np.zeros(3)) for c in range(4)]
        if not close([o["P"] for o in feet], s["foot_positions"]):
            bad.append(("anymal", k, "foot_positions"))
        if any(active):
            Jc = np.vstack([feet[c]["dCda"] for c in range(4) if active[c]])
            if not close(Jc, s["contact_jacobian"]):
                bad.append(("anymal", k, "contact_jacobian"))
            inv, info = fb.mjtjinv(np.array(s["dtau_da"]), np.array(s["contact_jacobian"]))
            if info != 0 or not close(inv, s["MJtJinv"], rtol=1e-8, atol=1e-8):
                bad.append(("anymal", k, "MJtJinv"))
    return bad


def _fixed_base_oracle(O, key):
    """(solver, iterations, first stage, last stage, last stage with controls) of a problem key of upstream_solvers.json."""
    if key.startswith("unocp_benchmark") or key.startswith("unparnmpc_benchmark"):
        prob, par = O.benchmark_problem(), key.startswith("unparnmpc")
        it = 20 if par else 50
    elif key == "config_space_ocp":
        prob, par, it = O.config_space_problem(), False, 30
    else:
        prob, par, it = O.task_space_problem(), key.endswith("unparnmpc"), 30
    s = (O.UnParNMPCSolver if par else O.UnOCPSolver)(prob)
    if prob.task_enabled:
        s.set_task_ref(O.task_ref_table(O.task_space_ref, 0.0, prob.T, prob.N, "unparnmpc" if par else "unocp"))
    return s, par, it


def run_fixed_base(O, key, q0, v0):
    s, par, iters = _fixed_base_oracle(O, key)
    q0, v0 = np.array(q0), np.array(v0)
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    if par:
        s.init_backward_correction(0.0)
    s.compute_kkt_residual(0.0, q0, v0)
    kkt, iterates = [s.kkt_error()], {}
    for it in range(1, iters + 1):
        s.update_solution(0.0, q0, v0, False)
        s.compute_kkt_residual(0.0, q0, v0)
        kkt.append(s.kkt_error())
        if it in (1, 2, iters):
            sol = {n: s.get_solution(n) for n in ("q", "v", "lmd", "gmm", "a", "u", "beta")}
            stages = []
            for i in range(len(sol["q"])):
                st = {n: sol[n][i].tolist() for n in ("q", "v", "lmd", "gmm")}
                if i < len(sol["u"]):
                    st.update({n: sol[n][i].tolist() for n in ("a", "u", "beta")})
                stages.append(st)
            iterates[str(it)] = stages
    return {"q0": q0.tolist(), "v0": v0.tolist(), "kkt": kkt, "iterates": iterates}


def check_fixed_base(rec, O):
    bad = []
    for key, ref in rec.items():
        if not isinstance(ref, dict) or "iterates" not in ref:
            continue
        got = run_fixed_base(O, key, ref["q0"], ref["v0"])
        for it, (a, b) in enumerate(zip(got["kkt"], ref["kkt"])):
            if not abs(a - b) <= KKT_ATOL + RTOL * abs(b):
                bad.append((key, "kkt", it, a, b))
                break
        if len(got["kkt"]) != len(ref["kkt"]):
            bad.append((key, "kkt length"))
        for it, stages in ref["iterates"].items():
            for i, st in enumerate(stages):
                for name, val in st.items():
                    if not close(got["iterates"][it][i][name], val, atol=1e-9):
                        bad.append((key, "iterate", it, i, name))
    return bad


def run_anymal_trotting(fb):
    import anymal_problems as ap
    pr = ap.TrottingProblem()
    o = pr.make_oracle(fb)
    o.compute_kkt_residual(0.0, pr.q0, pr.v0)
    kkt = [o.kkt_error()]
    for _ in range(25):
        o.update_solution(0.0, pr.q0, pr.v0)
        o.compute_kkt_residual(0.0, pr.q0, pr.v0)
        kkt.append(o.kkt_error())
    chain = o.chain()
    grid = [e for e, c in enumerate(chain) if c["kind"] in (fb.K_GRID, fb.K_TERMINAL)]
    final = {}
    for stage in (0, 11, 20, pr.N):
        e = grid[stage]
        final[str(stage)] = {"q": o.get(e, "q").tolist(), "v": o.get(e, "v")[:18].tolist()}
        if stage < pr.N:
            final[str(stage)]["u"] = o.get(e, "u").tolist()
    return {"kkt": kkt, "final": final}


def check_anymal(rec, fb):
    if "anymal_trotting" not in rec:
        return []
    ref, got, bad = rec["anymal_trotting"], run_anymal_trotting(fb), []
    for it, (a, b) in enumerate(zip(got["kkt"], ref["kkt"])):
        if not abs(a - b) <= KKT_ATOL + RTOL * abs(b):
            bad.append(("anymal_trotting", "kkt", it, a, b))
            break
    for stage, st in ref["final"].items():
        for name in ("q", "v", "u"):
            if name in st and not close(got["final"][stage][name], st[name], atol=1e-9):
                bad.append(("anymal_trotting", "final", stage, name))
    return bad


# ---------------------------------------------------------------------------------------------------------------
# the pins
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


@pytest.mark.skipif(not os.path.exists(ROBOT_FILE), reason=MISSING)
def test_upstream_robot_pins(oracle, fb):
    with open(ROBOT_FILE) as f:
        rec = json.load(f)
    assert "idocp" in rec["source"]
    assert check_robot(rec, oracle, fb) == []


@pytest.mark.skipif(not os.path.exists(SOLVER_FILE), reason=MISSING)
def test_upstream_solver_pins(oracle, fb):
    with open(SOLVER_FILE) as f:
        rec = json.load(f)
    assert check_fixed_base(rec, oracle) == []
    assert check_anymal(rec, fb) == []


def test_pin_status_is_reported():
    """The state of the pins is part of the test output either way (pytest -rs shows the skip reason)."""
    pinned = os.path.exists(ROBOT_FILE) and os.path.exists(SOLVER_FILE)
    print("oracle pinned to upstream idocp: %s" % ("YES (tests/golden/upstream_*.json)" if pinned else "NO -- " + MISSING))


def test_pin_consumer_on_oracle_made_files(oracle, fb):
    """Schema / plumbing test of the checkers above: files of the pin_dump schema made FROM THE ORACLE pass, and a perturbed
    value is caught.  (This pins nothing; it keeps the kit's consumer from rotting while the upstream files are absent.)"""
    rng = np.random.default_rng(5)
    robot = {"source": "oracle self-test", "iiwa14": [], "anymal": []}
    for _ in range(2):
        q, v, a = rng.uniform(-2, 2, 7), rng.uniform(-1.5, 1.5, 7), rng.uniform(-3, 3, 7)
        dq, dv, da = oracle.rnea_derivatives(q, v, a)
        R, p, J = oracle.frame_kinematics(q)
        robot["iiwa14"].append({"q": q.tolist(), "v": v.tolist(), "a": a.tolist(), "tau": oracle.rnea(q, v, a).tolist(),
                                "dtau_dq": dq.tolist(), "dtau_dv": dv.tolist(), "dtau_da": da.tolist(),
                                "frame22_position": p.tolist(), "frame22_rotation": R.tolist(), "frame22_jacobian_local": J.tolist()})
    for active in ([1, 0, 0, 1], [0, 0, 0, 0]):
        quat = rng.normal(size=4)
        q = np.concatenate([rng.uniform(-0.3, 0.3, 2), [0.5], quat / np.linalg.norm(quat), rng.uniform(-1, 1, 12)])
        v, a, f = rng.uniform(-1, 1, 18), rng.uniform(-2, 2, 18), rng.uniform(-20, 20, 12)
        fa = f.reshape(4, 3) * np.array(active, dtype=float)[:, None]
        tau, dq, dv, M = fb.rnea(q, v, a, fa.ravel(), derivatives=True)
        feet = [fb.contact(q, v, a, c, 0.05, np.zeros(3)) for c in range(4)]
        s = {"q": q.tolist(), "v": v.tolist(), "a": a.tolist(), "active": active, "f": f.tolist(), "tau": tau.tolist(),
             "dtau_dq": dq.tolist(), "dtau_dv": dv.tolist(), "dtau_da": M.tolist(), "foot_positions": [o["P"].tolist() for o in feet]}
        if any(active):
            Jc = np.vstack([feet[c]["dCda"] for c in range(4) if active[c]])
            s["contact_jacobian"] = Jc.tolist()
            s["MJtJinv"] = fb.mjtjinv(M, Jc)[0].tolist()
        robot["anymal"].append(s)
    assert check_robot(robot, oracle, fb) == []
    robot["iiwa14"][0]["tau"][3] *= 1 + 1e-7
    assert check_robot(robot, oracle, fb) == [("iiwa14", 0, "tau")]
    q0, v0 = [2.0] * 7, [0.0] * 7
    solvers = {"unocp_benchmark_reference_instance": run_fixed_base(oracle, "unocp_benchmark_reference_instance", q0, v0),
               "unparnmpc_benchmark_reference_instance": run_fixed_base(oracle, "unparnmpc_benchmark_reference_instance", q0, v0),
               "anymal_trotting": run_anymal_trotting(fb)}
    assert check_fixed_base(solvers, oracle) == []
    assert check_anymal(solvers, fb) == []
    solvers["unocp_benchmark_reference_instance"]["kkt"][7] *= 1 + 1e-6
    assert [b[:3] for b in check_fixed_base(solvers, oracle)] == [("unocp_benchmark_reference_instance", "kkt", 7)]

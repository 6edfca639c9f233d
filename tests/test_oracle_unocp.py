"""Oracle (CPU restatement of idocp's UnOCPSolver) against the identities that the reference's own
unit tests assert, against an independent numpy re-derivation of one stage, and against the
committed golden vectors."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _golden():
    with open(os.path.join(GOLDEN, "unocp_golden.json")) as f:
        return json.load(f)


def _numpy_stage(O, p, s, sn, slack, dual, stage):
    """Independent numpy re-derivation of SplitUnOCP::linearizeOCP for one stage (SURVEY.md A.2;
    mirrors test/unocp/split_unocp_test.cpp:92-137 and unconstrained_dynamics_test.cpp:68-126)."""
    dt = p.T / p.N
    g = lambda name: np.array(list(getattr(p, name)))
    q, v, a, u, beta, lmd, gmm = (s[k] for k in ("q", "v", "a", "u", "beta", "lmd", "gmm"))
    lq = dt * g("q_weight") * (q - g("q_ref"))
    lv = dt * g("v_weight") * (v - g("v_ref"))
    la = dt * g("a_weight") * a
    lu = dt * g("u_weight") * (u - g("u_ref"))
    Qqq = np.diag(dt * g("q_weight")); Qvv = dt * g("v_weight"); Qaa = dt * g("a_weight"); Quu = dt * g("u_weight")
    eps = p.barrier
    lims = [(q, g("q_min"), -1), (q, g("q_max"), +1), (v, -g("v_max"), -1), (v, g("v_max"), +1),
            (u, -g("u_max"), -1), (u, g("u_max"), +1)]
    active = [stage >= 2, stage >= 2, stage >= 1, stage >= 1, True, True]
    grads = [lq, lq, lv, lv, lu, lu]
    for c, ((x, lim, sg), act) in enumerate(zip(lims, active)):
        if not act:
            continue
        sl, z = slack[c], dual[c]
        r = (lim - x + sl) if sg < 0 else (x - lim + sl)
        cdual = sl * z - eps
        grads[c] += sg * dt * z                                  # augmentDualResidual
        grads[c] += sg * dt * (z * r - cdual) / sl               # condenseSlackAndDual
        h = dt * z / sl
        if c < 2: Qqq[np.diag_indices(7)] += h
        elif c < 4: Qvv = Qvv + h
        else: Quu = Quu + h
    Fq = q - sn["q"] + dt * v
    Fv = v + dt * a - sn["v"]
    lq += sn["lmd"] - lmd
    lv += dt * sn["lmd"] + sn["gmm"] - gmm
    la += dt * sn["gmm"]
    ID = O.rnea(q, v, a) - u
    dq, dv, da = O.rnea_derivatives(q, v, a)
    lq += dt * dq.T @ beta; lv += dt * dv.T @ beta; la += dt * da.T @ beta; lu -= dt * beta
    luc = lu + Quu * ID
    D = np.diag(Quu)
    Q = np.zeros((21, 21))
    Q[0:7, 0:7] = da.T @ D @ da + np.diag(Qaa)
    Q[0:7, 7:14] = da.T @ D @ dq
    Q[0:7, 14:21] = da.T @ D @ dv
    Q[7:14, 7:14] = dq.T @ D @ dq + Qqq
    Q[7:14, 14:21] = dq.T @ D @ dv
    Q[14:21, 14:21] = dv.T @ D @ dv + np.diag(Qvv)
    res = np.concatenate([Fq, Fv, la + da.T @ luc, lq + dq.T @ luc, lv + dv.T @ luc])
    return Q, res


def test_stage_linearisation_matches_numpy_rederivation(oracle):
    O = oracle
    p = O.benchmark_problem()
    rng = np.random.default_rng(3)
    s = O.UnOCPSolver(p)
    q0, v0 = rng.uniform(-1.5, 1.5, 7), rng.uniform(-0.5, 0.5, 7)
    s.set_solution("q", q0); s.set_solution("v", v0)
    for _ in range(3):   # move away from the trivial initial guess
        s.update_solution(0.0, q0, v0)
    sol = {n: s.get_solution(n) for n in ("q", "v", "a", "u", "beta", "lmd", "gmm")}
    slack, dual = s.get_constraint_data("slack"), s.get_constraint_data("dual")
    # a fresh solver restarted from that iterate, so get_unkkt shows the pure linearisation of
    # the LAST stage (its Riccati update only adds the diagonal terminal P)
    s.update_solution(0.0, q0, v0)
    for stage in (0, 1, 2, 7):
        cur = {k: (sol[k][stage] if stage < len(sol[k]) else None) for k in sol}
        nxt = {k: sol[k][stage + 1] for k in ("q", "v", "lmd", "gmm")}
        Q, res = _numpy_stage(O, p, cur, nxt, slack[stage], dual[stage], stage)
        Qo, ro = s.get_unkkt(stage)
        # Fx and the blocks are touched by the Riccati sweep afterwards (Q += A^T P A ...): compare the
        # parts it leaves alone: Fq, Fv, lq, lv exactly as linearised
        assert np.allclose(res[:14], ro[:14], rtol=1e-12, atol=1e-12)
        assert np.allclose(res[21:], ro[21:], rtol=1e-11, atol=1e-10)


def test_riccati_identities(oracle):
    """K = -Qaa^-1 [Qaq Qav], k = -Qaa^-1 la, P = Qxx - K^T Qaa K (symmetrised), terminal P
    (test/unocp/split_unriccati_factorizer_test.cpp:65-145; unriccati_recursion_test.cpp:74-113)."""
    O = oracle
    p = O.benchmark_problem()
    s = O.UnOCPSolver(p)
    q0 = np.full(7, 1.0); v0 = np.zeros(7)
    s.set_solution("q", q0); s.set_solution("v", v0)
    s.update_solution(0.0, q0, v0)
    rN = s.get_riccati(p.N)
    assert np.allclose(rN["Pqq"], np.diag([10.0] * 7)) and np.allclose(rN["Pvv"], np.diag([0.1] * 7))
    assert np.all(rN["Pqv"] == 0)
    for stage in (19, 10, 0):
        Q, res = s.get_unkkt(stage)          # blocks AFTER factorizeKKTMatrix (in place, like the reference)
        r = s.get_riccati(stage)
        Qaa, Qaq, Qav = Q[0:7, 0:7], Q[0:7, 7:14], Q[0:7, 14:21]
        Qqq, Qqv, Qvv = Q[7:14, 7:14], Q[7:14, 14:21], Q[14:21, 14:21]
        Qaa_l = np.tril(Qaa) + np.tril(Qaa, -1).T       # LLT reads the lower triangle
        K = -np.linalg.solve(Qaa_l, np.hstack([Qaq, Qav]))
        assert np.allclose(r["K"], K, rtol=1e-9, atol=1e-10)
        assert np.allclose(r["k"], -np.linalg.solve(Qaa_l, res[14:21]), rtol=1e-9, atol=1e-10)
        Kq, Kv = K[:, :7], K[:, 7:]
        Pqq = Qqq - Kq.T @ Qaa @ Kq
        Pvv = Qvv - Kv.T @ Qaa @ Kv
        assert np.allclose(r["Pqq"], 0.5 * (Pqq + Pqq.T), rtol=1e-9, atol=1e-9)
        assert np.allclose(r["Pqv"], Qqv - Kq.T @ Qaa @ Kv, rtol=1e-9, atol=1e-9)
        assert np.allclose(r["Pvv"], 0.5 * (Pvv + Pvv.T), rtol=1e-9, atol=1e-9)
        assert np.array_equal(r["Pqq"], r["Pqq"].T) and np.array_equal(r["Pvv"], r["Pvv"].T)
        # the Riccati matrix of a strictly convex stage problem is positive definite
        P = np.block([[r["Pqq"], r["Pqv"]], [r["Pqv"].T, r["Pvv"]]])
        assert np.all(np.linalg.eigvalsh(0.5 * (P + P.T)) > 0)


def test_newton_direction_solves_the_kkt_system(oracle):
    """The direction returned by the Riccati recursion satisfies the linearised dynamics
    dx+ = Fx + A dx + B da with dx0 = x0 - (q0, v0) (unocp_solver.cpp:100-101)."""
    O = oracle
    p = O.benchmark_problem()
    s = O.UnOCPSolver(p)
    rng = np.random.default_rng(5)
    x0q, x0v = rng.uniform(-1, 1, 7), rng.uniform(-0.3, 0.3, 7)
    s.set_solution("q", rng.uniform(-1, 1, 7)); s.set_solution("v", rng.uniform(-0.3, 0.3, 7))
    sol_q, sol_v, sol_a = s.get_solution("q"), s.get_solution("v"), s.get_solution("a")
    s.update_solution(0.0, x0q, x0v)
    dq, dv, da = s.get_direction("dq"), s.get_direction("dv"), s.get_direction("da")
    dt = p.T / p.N
    assert np.allclose(dq[0], x0q - sol_q[0]) and np.allclose(dv[0], x0v - sol_v[0])
    for i in range(p.N):
        Fq = sol_q[i] - sol_q[i + 1] + dt * sol_v[i]
        Fv = sol_v[i] + dt * sol_a[i] - sol_v[i + 1]
        assert np.allclose(dq[i + 1], Fq + dq[i] + dt * dv[i], rtol=1e-12, atol=1e-12)
        assert np.allclose(dv[i + 1], Fv + dv[i] + dt * da[i], rtol=1e-12, atol=1e-12)


def test_pdipm_rules(oracle):
    """slack initialisation, duality and fraction-to-boundary semantics (test/constraints/pdipm_test.cpp)."""
    O = oracle
    p = O.benchmark_problem()
    s = O.UnOCPSolver(p)
    q = np.array(list(p.q_max)) + 0.3      # violates the upper position limit
    s.set_solution("q", q)
    slack, dual = s.get_constraint_data("slack"), s.get_constraint_data("dual")
    assert np.all(slack[0, :4] == 0) and np.all(slack[1, :2] == 0)       # stage masks (constraints_data.hpp:18-43)
    assert np.all(slack[0, 4:] > 0) and np.all(slack[1, 2:] > 0) and np.all(slack[2] > 0)
    assert np.all(slack[2:] >= p.barrier)                                 # while (slack < eps) slack += eps
    assert np.allclose(slack[2:] * dual[2:], p.barrier)                   # dual = eps / slack
    assert not s.is_feasible()
    s.set_solution("q", np.zeros(7))
    assert s.is_feasible()
    s.update_solution(0.0, np.zeros(7), np.zeros(7))
    st = s.step_sizes()
    assert 0 < st[0] <= 1 and 0 < st[1] <= 1
    # after the update all slacks and duals stay strictly positive (fraction-to-boundary 0.995)
    assert np.all(s.get_constraint_data("slack")[2:] > 0) and np.all(s.get_constraint_data("dual")[2:] > 0)


def test_benchmark_instance_converges_like_the_example(oracle):
    """examples/iiwa14/unocp_benchmark.cpp: q = 2, v = 0, 50 iterations."""
    O = oracle
    p = O.benchmark_problem()
    s = O.UnOCPSolver(p)
    q0, v0 = np.full(7, 2.0), np.zeros(7)
    s.set_solution("q", q0); s.set_solution("v", v0)
    s.compute_kkt_residual(0.0, q0, v0)
    k0 = s.kkt_error()
    for _ in range(50):
        s.update_solution(0.0, q0, v0)
    s.compute_kkt_residual(0.0, q0, v0)
    assert s.kkt_error() < 1e-10 < k0
    assert s.is_feasible()


def test_golden_vectors(oracle):
    O = oracle
    G = _golden()
    for rec, prob, iters in [(G["unocp_benchmark_reference_instance"], O.benchmark_problem(), 50),
                             (G["config_space_ocp"], O.config_space_problem(), 30)]:
        s = O.UnOCPSolver(prob)
        q0, v0 = np.array(rec["q0"]), np.array(rec["v0"])
        s.set_solution("q", q0); s.set_solution("v", v0)
        s.compute_kkt_residual(0.0, q0, v0)
        kkt = [s.kkt_error()]
        for it in range(iters):
            s.update_solution(0.0, q0, v0)
            if str(it) in rec["directions"]:
                for n, ref in rec["directions"][str(it)].items():
                    assert np.array_equal(s.get_direction(n), np.array(ref)), n
            st = s.step_sizes()
            assert st[0] == rec["primal"][it] and st[1] == rec["dual"][it]
            s.compute_kkt_residual(0.0, q0, v0)
            kkt.append(s.kkt_error())
        assert np.array_equal(kkt, rec["kkt"])
        for n, ref in rec["final"].items():
            assert np.array_equal(s.get_solution(n), np.array(ref)), n
    for rec in G["rnea"]:
        q, v, a = np.array(rec["q"]), np.array(rec["v"]), np.array(rec["a"])
        assert np.allclose(O.rnea(q, v, a), rec["tau"], rtol=1e-12, atol=1e-12)   # libm sin/cos: not canonical
        dq, dv, da = O.rnea_derivatives(q, v, a)
        assert np.array_equal(dq, rec["dtau_dq"]) and np.array_equal(dv, rec["dtau_dv"])
        assert np.array_equal(da, rec["dtau_da"])


def test_filter_line_search_bounds(oracle):
    """test/line_search/unline_search_test.cpp:67-90: min_step <= alpha <= alpha_max."""
    O = oracle
    p = O.benchmark_problem()
    s = O.UnOCPSolver(p)
    q0, v0 = np.full(7, 2.0), np.zeros(7)
    s.set_solution("q", q0); s.set_solution("v", v0)
    for _ in range(5):
        s.update_solution(0.0, q0, v0, True)
        st = s.step_sizes()
        assert 0.05 <= st[0] <= st[2] + 1e-15 or st[0] == 0.05


def test_stage_threads_do_not_change_results(oracle):
    """reference threading (OpenMP over stages, unocp_benchmark.cpp:42) is a pure scheduling choice."""
    O = oracle
    p = O.benchmark_problem()
    a, b = O.UnOCPSolver(p), O.UnOCPSolver(p)
    b.set_stage_threads(4)
    q0, v0 = np.full(7, 1.5), np.zeros(7)
    for s in (a, b):
        s.set_solution("q", q0); s.set_solution("v", v0)
        for _ in range(4):
            s.update_solution(0.0, q0, v0)
    assert np.array_equal(a.get_solution("q"), b.get_solution("q"))
    assert np.array_equal(a.get_solution("u"), b.get_solution("u"))


def test_acceleration_limits_bind_and_converge(oracle):
    """JointAccelerationLowerLimit / UpperLimit (src/constraints/joint_acceleration_*_limit.cpp) on the unocp_benchmark problem:
    the iteration converges, the bound is active somewhere and respected everywhere, complementarity holds."""
    O = oracle
    free = O.UnOCPSolver(O.benchmark_problem())
    q0, v0 = np.full(7, 2.0), np.zeros(7)
    for s in (free,):
        s.set_solution("q", q0)
        s.set_solution("v", v0)
    for _ in range(60):
        free.update_solution(0.0, q0, v0)
    amax_free = np.abs(free.get_solution("a")).max()
    limit = 0.5 * amax_free
    p = O.benchmark_problem()
    p.enable_acc[0] = p.enable_acc[1] = 1
    for j in range(7):
        p.a_min[j], p.a_max[j] = -limit, limit
    s = O.UnOCPSolver(p)
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    assert np.all(s.get_constraint_data("acc_slack") > 0)
    for _ in range(80):
        s.update_solution(0.0, q0, v0)
    s.compute_kkt_residual(0.0, q0, v0)
    assert s.kkt_error() < 1e-8
    a = s.get_solution("a")
    assert np.abs(a).max() < limit and np.abs(a).max() > 0.98 * limit
    sl, du = s.get_constraint_data("acc_slack"), s.get_constraint_data("acc_dual")
    assert np.all(sl > 0) and np.all(du > 0)
    assert np.allclose(sl * du, p.barrier, rtol=1e-6)
    # slack = margin at the solution: a - amin, amax - a
    assert np.allclose(sl[:, 0], a + limit, atol=1e-9) and np.allclose(sl[:, 1], limit - a, atol=1e-9)

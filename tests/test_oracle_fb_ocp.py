"""Oracle of the ANYmal OCPSolver path (oracle/fb_ocp.c): identities of the reference's unit tests and convergence of
its shipped example.  The reference pins no numbers here -- "parity unpinned"."""
import numpy as np
import pytest

import anymal_problems as ap


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


def run(ocp, pr, iters, t=0.0, q=None, v=None):
    q = pr.q0 if q is None else q
    v = pr.v0 if v is None else v
    hist = []
    for _ in range(iters):
        ocp.compute_kkt_residual(t, q, v)
        hist.append(ocp.kkt_error())
        assert ocp.update_solution(t, q, v) == 0
    ocp.compute_kkt_residual(t, q, v)
    hist.append(ocp.kkt_error())
    return np.array(hist)


def test_chain_of_the_trotting_example(fb):
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    ch = ocp.chain()
    kinds = [c["kind"] for c in ch]
    # N = 30 grid stages + terminal, one lift (t = 0.5), two impulses (t = 1.0, 1.5) with their aux stages
    assert len(ch) == 30 + 1 + 1 + 2 * 2
    assert kinds.count(fb.K_LIFT) == 1 and kinds.count(fb.K_IMPULSE) == 2 and kinds.count(fb.K_AUX) == 2
    # the switching constraint sits two chain elements ahead of every impulse (ocp_linearizer.hxx:139-150)
    for e, c in enumerate(ch):
        expect = 6 if (e + 2 < len(ch) and ch[e + 2]["kind"] == fb.K_IMPULSE) else 0
        assert c["dimi"] == expect
    assert abs(sum(c["dt"] for c in ch) - pr.T) < 1e-12


def test_trotting_example_converges(fb):
    # examples/anymal/anymal_trotting.cpp: ocpbenchmarker::Convergence(ocp_solver, t, q, v, 25, false)
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    hist = run(ocp, pr, 25)
    assert hist[0] > 50 and hist[-1] < 1e-9
    assert np.all(np.isfinite(hist))
    ch = ocp.chain()
    qT = ocp.get(len(ch) - 1, "q")
    assert 0.05 < qT[0] < 0.40 and abs(qT[2] - 0.4792) < 0.05     # walked forward, still upright
    for e, c in enumerate(ch[:-1]):
        assert np.abs(ocp.get(e, "Fq")).max() < 1e-9 and np.abs(ocp.get(e, "Fv")).max() < 1e-9
        assert np.abs(ocp.get(e, "IDC")[:18 + c["dimf"]]).max() < 1e-8
        assert np.abs(ocp.get(e, "P")[:c["dimi"]]).max(initial=0.0) < 1e-9
        f = ocp.get(e, "f").reshape(4, 3)
        if c["kind"] != fb.K_IMPULSE:
            assert np.abs(ocp.get(e, "u")).max() <= 80.0
    # friction cone on the active feet of a mid-stance stage
    f = ocp.get(3, "f").reshape(4, 3)
    assert np.all(f[:, 2] > 0) and np.all(np.abs(f[:, :2]) <= 0.7 / np.sqrt(2) * f[:, 2:3] + 1e-9)
    assert abs(f[:, 2].sum() - ap.TOTAL_WEIGHT) < 0.2 * ap.TOTAL_WEIGHT


def test_expanded_direction_solves_the_linearised_contact_dynamics(fb):
    # test/ocp/contact_dynamics_test.cpp: the condensed/expanded (da, df) satisfy the linearised ID and Baumgarte rows
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    run(ocp, pr, 2)
    # one more update: every stage keeps the linearisation data of the point the direction was computed at
    ocp.update_solution(0.0, pr.q0, pr.v0)
    for e, c in enumerate(ocp.chain()[:-1]):
        dimf = c["dimf"]
        n = 18 + dimf
        dIDC = ocp.get(e, "dIDCdqv").reshape(30, 36)[:n]
        M = ocp.get(e, "Mm").reshape(18, 18)
        J = ocp.get(e, "dCda").reshape(12, 18)[:dimf]
        IDC = ocp.get(e, "IDC")[:n]
        dx = np.concatenate([ocp.get(e, "dq"), ocp.get(e, "dv")])
        daf = ocp.get(e, "daf")[:n]
        da, df = daf[:18], daf[18:]
        du_full = np.zeros(18)
        if c["kind"] != fb.K_IMPULSE:
            du_full[6:] = ocp.get(e, "du")
        lhs_id = dIDC[:18] @ dx + M @ da - J.T @ df - du_full + IDC[:18]
        lhs_c = dIDC[18:] @ dx + J @ da + IDC[18:]
        scale = max(1.0, np.abs(IDC).max())
        assert np.abs(lhs_id).max() < 1e-8 * scale, (e, c, np.abs(lhs_id).max())
        assert np.abs(lhs_c).max() < 1e-8 * scale, (e, c)
        K = np.zeros((n, n))
        K[:18, :18] = M
        K[:18, 18:] = J.T
        K[18:, :18] = J
        assert np.allclose(ocp.get(e, "MJtJinv").reshape(30, 30)[:n, :n] @ K, np.eye(n), atol=1e-8)


def test_riccati_factorisation_identities(fb):
    # test/ocp/split_riccati_factorizer_test.cpp: P symmetric, K = -G^-1 H^T on unconstrained stages,
    # forward recursion reproduces the linearised state equation
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    run(ocp, pr, 1)
    ocp.update_solution(0.0, pr.q0, pr.v0)
    ch = ocp.chain()
    for e, c in enumerate(ch[:-1]):
        Pqq, Pvv = ocp.get(e, "Pqq").reshape(18, 18), ocp.get(e, "Pvv").reshape(18, 18)
        if c["dimi"] == 0:
            assert np.array_equal(Pqq, Pqq.T) and np.array_equal(Pvv, Pvv.T)
        else:   # the two Schur corrections are subtracted after the symmetrisation (split_riccati_factorizer.hxx:88-96)
            assert np.allclose(Pqq, Pqq.T, rtol=1e-12, atol=1e-9) and np.allclose(Pvv, Pvv.T, rtol=1e-12, atol=1e-9)
        if c["kind"] == fb.K_IMPULSE:
            continue
        G = ocp.get(e, "Quu").reshape(18, 18)[6:, 6:]
        H = ocp.get(e, "Qxu").reshape(36, 18)[:, 6:]
        K = ocp.get(e, "K").reshape(12, 36)
        assert np.all(np.linalg.eigvalsh(0.5 * (G + G.T)) > 0)
        if c["dimi"] == 0:
            assert np.allclose(G @ K, -H.T, rtol=1e-9, atol=1e-9 * np.abs(H).max())
        else:
            # constrained stage (split_riccati_factorizer.hxx:55-100): [G D^T; D 0] [K; M] = -[H^T; Phix]
            D = ocp.get(e, "Phiu").reshape(12, 12)[:c["dimi"]]
            Mx = ocp.get(e, "cM").reshape(12, 36)[:c["dimi"]]
            Phix = ocp.get(e, "Phix").reshape(12, 36)[:c["dimi"]]
            assert np.allclose(G @ K + D.T @ Mx, -H.T, rtol=1e-8, atol=1e-8 * np.abs(H).max())
            assert np.allclose(D @ K, -Phix, rtol=1e-8, atol=1e-8 * max(1, np.abs(Phix).max()))
            # the direction keeps the linearised switching constraint: Phix dx + Phiu du + P = 0
            dx = np.concatenate([ocp.get(e, "dq"), ocp.get(e, "dv")])
            r = Phix @ dx + D @ ocp.get(e, "du") + ocp.get(e, "P")[:c["dimi"]]
            assert np.abs(r).max() < 1e-9


def test_threads_do_not_change_bits(fb):
    pr = ap.TrottingProblem()
    a, b = pr.make_oracle(fb), pr.make_oracle(fb)
    b.set_threads(4)
    ha, hb = run(a, pr, 4), run(b, pr, 4)
    assert np.array_equal(ha, hb)
    for e in range(len(a.chain())):
        assert np.array_equal(a.get(e, "q"), b.get(e, "q")) and np.array_equal(a.get(e, "lmd"), b.get(e, "lmd"))


def test_four_step_trot_and_perturbed_initial_state(fb):
    pr = ap.TrottingProblem(steps=4)
    rng = np.random.default_rng(0)
    dq = np.concatenate([rng.uniform(-0.01, 0.01, 3), rng.uniform(-0.02, 0.02, 3), rng.uniform(-0.02, 0.02, 12)])
    q0 = fb.integrate(pr.q0, dq)
    v0 = rng.uniform(-0.1, 0.1, 18)
    ocp = pr.make_oracle(fb, q0=q0, v0=v0)
    hist = run(ocp, pr, 40, q=q0, v=v0)
    assert hist[-1] < 1e-8, hist[-5:]


def test_flight_phase_lift_stage_and_four_foot_touch_down(fb):
    # a flight phase (dimf = 0), a lift stage carrying the switching constraint of a 12-dimensional touch-down.
    # Gauss-Newton without line search converges slowly here, as upstream does (anymal_jumping.cpp: 155 iterations).
    pr = ap.JumpingProblem(0.1, 0.6, 0.75, 1.3, 26)
    for nm, w in (("q_weight", [1, 1, 1] + [10] * 15), ("v_weight", [0.01] * 3 + [0.1] * 15), ("a_weight", [0.01] * 18)):
        pr.problem.set(nm, w)
        pr.problem.set({"q_weight": "qf_weight", "v_weight": "vf_weight", "a_weight": "dvi_weight"}[nm], w)
        if nm != "a_weight":
            pr.problem.set({"q_weight": "qi_weight", "v_weight": "vi_weight"}[nm], w)
    ocp = pr.make_oracle(fb)
    ch = ocp.chain()
    kinds = [c["kind"] for c in ch]
    assert kinds.count(fb.K_LIFT) == 1 and kinds.count(fb.K_IMPULSE) == 1
    assert any(c["dimf"] == 0 for c in ch)
    lift = kinds.index(fb.K_LIFT)
    assert sum(c["dimi"] == 12 for c in ch) == 1
    hist = run(ocp, pr, 80)
    assert np.all(np.isfinite(hist)) and hist[-1] < 2e-2 and hist[-1] < hist[-10] < hist[-20]
    # the feet are off the ground during the flight and land on the shifted contact points
    imp = kinds.index(fb.K_IMPULSE)
    P = np.stack([fb.contact(ocp.get(imp, "q"), np.zeros(18), np.zeros(18), i, 0.05, np.zeros(3))["P"] for i in range(4)])
    assert np.allclose(P[:, 0], ap.standing_contact_points(fb)[:, 0] + 0.1, atol=5e-3)


def test_filter_line_search(fb):
    # LineSearch::computeStepSize (line_search.hpp:62-93) + LineSearchFilter (line_search_filter.cpp:34-65)
    pr = ap.JumpingProblem(0.1, 0.6, 0.75, 1.3, 26)
    a, b = pr.make_oracle(fb), pr.make_oracle(fb)
    for it in range(6):
        # the direction and the fraction-to-boundary step do not depend on the flag
        assert b.update_solution(0.0, pr.q0, pr.v0, False) == 0
        amax = b.step_sizes()[0]
        n0 = a.filter_size()
        assert a.update_solution(0.0, pr.q0, pr.v0, True) == 0
        alpha = a.step_sizes()[0]
        # alpha is alpha_max * 0.75^k for some k >= 0, or the floor 0.05
        ks = [amax * 0.75 ** k for k in range(12)]
        assert alpha == 0.05 or any(alpha == x for x in ks), (alpha, amax)
        assert a.filter_size() >= 1 and a.filter_size() <= n0 + 2
        # keep the two solvers on the same iterate: copy a's iterate into b
        for e in range(len(a.chain())):
            for nm in ("q", "v", "a", "u", "f", "lmd", "gmm", "beta", "mu", "nu_passive", "xi"):
                b.set(e, nm, a.get(e, nm))
        # (slack / dual are not settable: stop comparing directions after the first divergence)
        break
    # cost / violation of the current point at a converged solution: every stage is feasible EXCEPT the grid stages
    # right before an event -- upstream closes their state equation with the next GRID stage instead of the lift /
    # impulse stage (line_search.cpp:84-121: the first if / else-if is overwritten by the second if / else); kept
    pr = ap.TrottingProblem()
    c = pr.make_oracle(fb)
    for _ in range(25):
        c.update_solution(0.0, pr.q0, pr.v0, False)
    tot = c.cost_and_violation(0.0)
    ch = c.chain()
    per_stage = np.array([c.get(e, "ls_viol")[0] for e in range(len(ch))])
    before_event = [e for e in range(len(ch) - 1) if ch[e]["kind"] == fb.K_GRID and ch[e + 1]["kind"] in (fb.K_IMPULSE, fb.K_LIFT)]
    assert len(before_event) == 3
    mask = np.ones(len(ch), bool)
    mask[before_event] = False
    assert np.all(per_stage[mask] < 1e-8) and np.all(per_stage[before_event] > 0.5)
    assert abs(tot[1] - per_stage.sum()) < 1e-9 and np.isfinite(tot[0])
    c.update_solution(0.0, pr.q0, pr.v0, True)
    assert c.filter_size() >= 1
    c.clear_line_search_filter()
    assert c.filter_size() == 0


def test_running_gait_one_stride(fb):
    # examples/anymal/anymal_running.cpp shortened to one stride of the loop (flight phases between the swings)
    pr = ap.RunningProblem(steps=1)
    ocp = pr.make_oracle(fb)
    ocp.set_threads(8)
    ch = ocp.chain()
    kinds = [c["kind"] for c in ch]
    assert kinds.count(fb.K_IMPULSE) == 8 and kinds.count(fb.K_LIFT) == 5 and any(c["dimf"] == 0 for c in ch)
    hist = run(ocp, pr, 40)
    assert np.all(np.isfinite(hist)) and hist[-1] < 1e-5 * hist[0]
    full = ap.RunningProblem(steps=10)
    o2 = full.make_oracle(fb)
    k2 = [c["kind"] for c in o2.chain()]
    assert (full.T, full.N) == (7.0, 240) and k2.count(fb.K_IMPULSE) == 26 and k2.count(fb.K_LIFT) == 14   # the example's schedule


# ---- FrictionCone / ImpulseFrictionCone, JointAcceleration{Lower,Upper}Limit ----
SOL_FIELDS = ["q", "v", "a", "u", "f", "lmd", "gmm", "beta", "mu", "nu_passive", "xi"]


def cone_g(mu, f):
    return np.array([-f[2], f[0] ** 2 + f[1] ** 2 - mu ** 2 * f[2] ** 2])


def test_nonlinear_cone_and_acceleration_limit_gradients(fb):
    """augmentDualResidual of FrictionCone / ImpulseFrictionCone (friction_cone.cpp:100-118) and of the acceleration limits
    (joint_acceleration_lower_limit.cpp:52-56): lf and la of the same iterate with and without the components differ by
    dt J^T dual, J by central differences of the constraint functions."""
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem())
    ocp = pr.make_oracle(fb)
    run(ocp, pr, 6)
    pr0 = ap.TrottingProblem()
    pr0.problem.enable[6] = pr0.problem.enable[7] = 0
    bare = pr0.make_oracle(fb)
    ch = ocp.chain()
    for e in range(len(ch)):
        for nm in SOL_FIELDS:
            bare.set(e, nm, ocp.get(e, nm))
    bare.compute_kkt_residual(0.0, pr.q0, pr.v0)
    mu, checked = pr.problem.mu, 0
    for e, c in enumerate(ch[:-1]):
        imp = c["kind"] == fb.K_IMPULSE
        dt = 1.0 if imp else c["dt"]
        dual = ocp.get(e, "dual")
        f = ocp.get(e, "f").reshape(4, 3)
        cone = dual[92:112] if imp else dual[72:92]
        expect, k = np.zeros(12), 0
        for i in range(4):
            if not ocp.get(e, "active")[i]:
                continue
            J = np.zeros((2, 3))
            for x in range(3):
                h = 1e-4 * max(1.0, abs(f[i, x]))
                fp, fm = f[i].copy(), f[i].copy()
                fp[x] += h
                fm[x] -= h
                J[:, x] = (cone_g(mu, fp) - cone_g(mu, fm)) / (2 * h)
            expect[3 * k:3 * k + 3] = dt * (J.T @ cone[2 * i:2 * i + 2])
            k += 1
        got = ocp.get(e, "lf") - bare.get(e, "lf")
        assert np.allclose(got[:3 * k], expect[:3 * k], rtol=1e-6, atol=1e-9 * max(1.0, np.abs(ocp.get(e, "lf")).max())), (e, got, expect)
        assert np.all(cone[8:] == 0.0)        # two rows per contact: the rest of the component's storage is dead
        checked += k
        if not imp:
            dla = ocp.get(e, "la") - bare.get(e, "la")
            assert np.allclose(dla[6:], dt * (dual[124:136] - dual[112:124]), rtol=1e-9, atol=1e-12)
            assert np.all(dla[:6] == 0.0)
        else:
            assert np.all(dual[112:136] == 0.0)   # acceleration-level constraints do not exist at an impulse
    assert checked > 60


def test_nonlinear_cone_direction_is_the_newton_step_of_the_residual(fb):
    """computeSlackAndDualDirection (friction_cone.cpp:152-181, joint_acceleration_lower_limit.cpp:72-77): after the step
    alpha the linear rows satisfy residual+ = (1 - alpha) residual and the cone row
    residual+ = (1 - alpha) residual + alpha^2 g_quadratic(df)."""
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem())
    ocp = pr.make_oracle(fb)
    run(ocp, pr, 3)
    ch = ocp.chain()
    mu, amin, amax = pr.problem.mu, -9.0, 9.0

    def residuals(e, c):
        sl = ocp.get(e, "slack")
        f = ocp.get(e, "f").reshape(4, 3)
        a = ocp.get(e, "a")
        imp = c["kind"] == fb.K_IMPULSE
        cone = sl[92:112] if imp else sl[72:92]
        r = [cone_g(mu, f[i]) + cone[2 * i:2 * i + 2] if ocp.get(e, "active")[i] else np.zeros(2) for i in range(4)]
        racc = np.zeros(24) if imp else np.concatenate([amin - a[6:] + sl[112:124], a[6:] - amax + sl[124:136]])
        return np.array(r), racc, f
    before = [residuals(e, c) for e, c in enumerate(ch[:-1])]
    assert ocp.update_solution(0.0, pr.q0, pr.v0) == 0
    alpha = ocp.step_sizes()[0]
    assert 0 < alpha <= 1
    for e, c in enumerate(ch[:-1]):
        r0, a0, f0 = before[e]
        r1, a1, f1 = residuals(e, c)
        df = (f1 - f0) / alpha
        quad = df[:, 0] ** 2 + df[:, 1] ** 2 - mu ** 2 * df[:, 2] ** 2
        act = ocp.get(e, "active").astype(bool)
        scale = max(1.0, np.abs(r0).max())
        assert np.allclose(r1[act, 0], (1 - alpha) * r0[act, 0], atol=1e-9 * scale), e
        assert np.allclose(r1[act, 1], (1 - alpha) * r0[act, 1] + alpha ** 2 * quad[act], atol=1e-8 * scale), e
        assert np.allclose(a1, (1 - alpha) * a0, atol=1e-10 * max(1.0, np.abs(a0).max())), e


def test_nonlinear_cones_converge_to_a_feasible_trot(fb):
    # the reference drops the curvature of the cone from the Hessian (friction_cone.cpp:133-136 adds only r r^T), so the
    # iteration converges linearly, not quadratically
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem(), a_limit=None)
    ocp = pr.make_oracle(fb)
    hist = run(ocp, pr, 60)
    assert np.all(np.isfinite(hist)) and hist[-1] < 1e-8 * hist[0]
    mu = pr.problem.mu
    for e, c in enumerate(ocp.chain()[:-1]):
        f = ocp.get(e, "f").reshape(4, 3)
        sl, du = ocp.get(e, "slack"), ocp.get(e, "dual")
        o = 92 if c["kind"] == fb.K_IMPULSE else 72
        for i in range(4):
            if ocp.get(e, "active")[i]:
                assert f[i, 2] > 0 and f[i, 0] ** 2 + f[i, 1] ** 2 < mu ** 2 * f[i, 2] ** 2
                assert np.all(sl[o + 2 * i:o + 2 * i + 2] > 0) and np.all(du[o + 2 * i:o + 2 * i + 2] > 0)
                # complementarity: slack * dual = barrier at the solution
                assert np.allclose(sl[o + 2 * i:o + 2 * i + 2] * du[o + 2 * i:o + 2 * i + 2], pr.problem.barrier, rtol=1e-5)


def test_acceleration_limits_bind_and_converge(fb):
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem(), cones=False, a_limit=12.0)
    ocp = pr.make_oracle(fb)
    hist = run(ocp, pr, 40)
    assert np.all(np.isfinite(hist)) and hist[-1] < 1e-9
    amax = 0.0
    for e, c in enumerate(ocp.chain()[:-1]):
        if c["kind"] != fb.K_IMPULSE:
            amax = max(amax, np.abs(ocp.get(e, "a")[6:]).max())
            assert np.all(ocp.get(e, "slack")[112:136] > 0)
    assert 11.5 < amax < 12.0      # the limit binds (18 rad/s^2 without it)


def test_contact_distance_jacobian_and_convergence(fb):
    """ContactDistance (src/constraints/contact_distance.cpp).  Mode 2 (consistent): the rows J2 are the derivative of the height
    of the swing feet (central differences on the manifold), the trot converges and the swing feet stay above the ground.
    Mode 1 (the reference literally: row 2 of the LOCAL frame Jacobian, robot.hxx:182-188): same residuals, but J2 is NOT that
    derivative -- the shank-fixed foot frame is not level -- and the Newton iteration does not converge on the trot."""
    def foot_z(q, i):
        z = np.zeros(18)
        return fb.contact(q, z, z, i, 0.05, np.zeros(3))["P"][2]
    results = {}
    for mode in (1, 2):
        pr = ap.TrottingProblem()
        pr.problem.enable_distance = mode
        rng = np.random.default_rng(3)
        q0 = fb.integrate(pr.q0, rng.uniform(-0.05, 0.05, 18))
        ocp = pr.make_oracle(fb, q0=q0, v0=pr.v0)
        ocp.update_solution(0.0, q0, pr.v0)
        ocp.compute_kkt_residual(0.0, q0, pr.v0)
        e = next(e for e, c in enumerate(ocp.chain()) if c["kind"] == fb.K_GRID and c["index"] >= 12 and not all(ocp.get(e, "active")))
        q = ocp.get(e, "q")
        J = ocp.get(e, "cdJ").reshape(4, 18)
        act = ocp.get(e, "active")
        worst = 0.0
        for i in range(4):
            if act[i]:
                continue
            assert abs(ocp.get(e, "cdz")[i] - foot_z(q, i)) < 1e-14
            fd = np.zeros(18)
            for c in range(18):
                d = np.zeros(18)
                d[c] = 1e-6
                fd[c] = (foot_z(fb.integrate(q, d), i) - foot_z(fb.integrate(q, -d), i)) / 2e-6
            worst = max(worst, np.abs(J[i] - fd).max())
            # residual = -z + slack for the legs in the air, zero rows for the legs on the ground
            sl = ocp.get(e, "slack")[136:140]
            assert sl[i] > 0
        results[mode] = worst
        hist = run(ocp, pr, 30, q=q0)
        if mode == 2:
            assert hist[-1] < 1e-8
            for e2, c in enumerate(ocp.chain()[:-1]):
                if c["kind"] == fb.K_GRID and c["index"] >= 2:
                    a2 = ocp.get(e2, "active")
                    for i in range(4):
                        if not a2[i]:
                            assert foot_z(ocp.get(e2, "q"), i) > 0
        else:
            assert not hist[-1] < 1e-3      # the reference's rows: no convergence
    assert results[2] < 1e-7 and results[1] > 1e-2

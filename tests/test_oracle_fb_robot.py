"""Floating-base (ANYmal) rigid-body layer of the oracle vs independent formulations (the pinocchio boundary).

The reference pins nothing here (test/robot/robot_test.cpp compares against pinocchio itself, absent), so
oracle/fb_robot.h is validated against (i) a body-frame RNEA on the tree written independently in numpy
(oracle/fb_mirror.py), (ii) central finite differences on the configuration manifold, (iii) scipy expm/logm
for the SE(3) maps, (iv) the identities the reference's tests assert (robot_test.cpp:545-565 MJtJinv,
:141-219 configuration-space operators)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


@pytest.fixture(scope="module")
def fbm():
    import fb_mirror
    return fb_mirror


def rand_q(rng, big=False):
    q = np.zeros(19)
    q[:3] = rng.uniform(-1, 1, 3)
    quat = rng.normal(size=4)
    q[3:7] = quat / np.linalg.norm(quat)
    q[7:] = rng.uniform(-1.5, 1.5, 12)
    return q


def rand_state(rng):
    return rand_q(rng), rng.uniform(-2, 2, 18), rng.uniform(-5, 5, 18)


def fd_q(fb, fun, q, eps=1e-6):
    """Central differences of fun(q) along the local tangent directions (pinocchio's convention)."""
    cols = []
    for c in range(18):
        d = np.zeros(18)
        d[c] = eps
        cols.append((fun(fb.integrate(q, d)) - fun(fb.integrate(q, -d))) / (2 * eps))
    return np.stack(cols, axis=-1)


def fd_x(fun, x, eps=1e-6):
    cols = []
    for c in range(x.size):
        d = np.zeros(x.size)
        d[c] = eps
        cols.append((fun(x + d) - fun(x - d)) / (2 * eps))
    return np.stack(cols, axis=-1)


def test_model_facts(fbm):
    m = fbm.load_model()
    assert m["frames"][14] == "LF_FOOT" and m["frames"][24] == "LH_FOOT"
    assert m["frames"][34] == "RF_FOOT" and m["frames"][44] == "RH_FOOT"   # examples/anymal/anymal_trotting.cpp:30
    assert m["contact_parent"] == [2, 5, 8, 11]
    assert abs(m["total_mass"] - 30.4754) < 1e-3
    assert np.allclose(m["effort"], 80) and np.allclose(m["v_max"], 15) and np.allclose(m["q_max"], 9.42)


def test_exp_log_match_scipy(fb, fbm):
    rng = np.random.default_rng(0)
    for scale in (1e-6, 1e-3, 0.3, 1.2):
        for _ in range(5):
            nu = rng.normal(size=6) * scale
            if np.linalg.norm(nu[3:]) > 3.0:
                nu[3:] *= 3.0 / np.linalg.norm(nu[3:])
            R, p = fb.exp6(nu)
            R2, p2 = fbm.exp6(nu)
            assert np.allclose(R, R2, atol=1e-13) and np.allclose(p, p2, atol=1e-13)
            assert np.allclose(fb.log6(R, p), nu, atol=1e-9 * max(1, scale))


def test_integrate_subtract_roundtrip(fb, fbm):
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rand_q(rng)
        v = rng.normal(size=18) * 0.7
        q1 = fb.integrate(q, v, 0.8)
        assert abs(np.linalg.norm(q1[3:7]) - 1) < 1e-12
        assert np.allclose(q1, fbm.integrate(q, v, 0.8), atol=1e-12)
        assert np.allclose(fb.subtract(q1, q), 0.8 * v, atol=1e-11)


def test_dsubtract_and_dintegrate_by_finite_differences(fb):
    rng = np.random.default_rng(2)
    for _ in range(5):
        qp, qm = rand_q(rng), rand_q(rng)
        jp, jm = fb.dsubtract(qp, qm)
        fdp = fd_q(fb, lambda x: fb.subtract(x, qm), qp)
        fdm = fd_q(fb, lambda x: fb.subtract(qp, x), qm)
        assert np.allclose(jp, fdp[:6, :6], atol=2e-7)
        assert np.allclose(jm, fdm[:6, :6], atol=2e-7)
        assert np.allclose(fdp[6:, 6:], np.eye(12), atol=1e-8) and np.allclose(fdm[6:, 6:], -np.eye(12), atol=1e-8)
        inv = fb.dsubtract_inverse(jm)
        assert np.allclose(inv @ jm, np.eye(6), atol=1e-12)
        # dIntegrate: d(q (+) v)/dq and /dv expressed in the tangent at the result
        v = rng.normal(size=18) * 0.5
        jq, jv = fb.dintegrate(v[:6])
        q1 = fb.integrate(qp, v)
        fq = fd_q(fb, lambda x: fb.subtract(fb.integrate(x, v), q1), qp)
        fv = fd_x(lambda x: fb.subtract(fb.integrate(qp, x), q1), v)
        assert np.allclose(jq, fq[:6, :6], atol=2e-7)
        assert np.allclose(jv, fv[:6, :6], atol=2e-7)


def test_rnea_matches_body_frame_numpy(fb, fbm):
    m = fbm.load_model()
    rng = np.random.default_rng(3)
    for _ in range(10):
        q, v, a = rand_state(rng)
        f = rng.uniform(-50, 50, (4, 3))
        assert np.allclose(fb.rnea(q, v, a, f), fbm.rnea_body(m, q, v, a, f), rtol=1e-11, atol=1e-10)
        assert np.allclose(fb.rnea(q, 0 * v, a, f, gravity=0.0), fbm.rnea_body(m, q, 0 * v, a, f, gravity=0.0),
                           rtol=1e-11, atol=1e-10)


def test_standing_weight(fb, fbm):
    m = fbm.load_model()
    q = np.array([0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0])
    tau = fb.rnea(q, np.zeros(18), np.zeros(18))
    assert abs(tau[2] - m["total_mass"] * 9.81) < 1e-9 and abs(tau[0]) < 1e-12 and abs(tau[1]) < 1e-12


def test_rnea_derivatives_by_finite_differences(fb):
    rng = np.random.default_rng(4)
    for trial in range(4):
        q, v, a = rand_state(rng)
        f = rng.uniform(-50, 50, (4, 3))
        tau, dq, dv, M = fb.rnea(q, v, a, f, derivatives=True)
        assert np.array_equal(tau, fb.rnea(q, v, a, f))
        fq = fd_q(fb, lambda x: fb.rnea(x, v, a, f), q)
        fv = fd_x(lambda x: fb.rnea(q, x, a, f), v)
        fa = fd_x(lambda x: fb.rnea(q, v, x, f), a)
        sc = max(1.0, np.abs(fq).max())
        assert np.allclose(dq, fq, atol=5e-7 * sc), np.abs(dq - fq).max()
        assert np.allclose(dv, fv, atol=5e-7 * sc), np.abs(dv - fv).max()
        assert np.allclose(M, fa, atol=5e-7 * sc)
        assert np.array_equal(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
        # impulse model: zero gravity, zero velocity (Robot::RNEAImpulseDerivatives, robot.hxx:518-535)
        tau0, dq0, _, M0 = fb.rnea(q, 0 * v, a, f, gravity=0.0, derivatives=True)
        fq0 = fd_q(fb, lambda x: fb.rnea(x, 0 * v, a, f, gravity=0.0), q)
        assert np.allclose(dq0, fq0, atol=5e-7 * sc)
        assert np.allclose(M0, M, atol=1e-12)


def test_frame_kinematics_and_derivatives(fb, fbm):
    m = fbm.load_model()
    rng = np.random.default_rng(5)
    for trial in range(3):
        q, v, a = rand_state(rng)
        for i in range(4):
            cp = rng.uniform(-1, 1, 3)
            dt = 0.05
            o = fb.contact(q, v, a, i, dt, cp)
            P, R, vF, aF, acl = fbm.frame_state(m, q, v, a, i)
            assert np.allclose(o["P"], P, atol=1e-13) and np.allclose(o["vF"], vF, atol=1e-12)
            assert np.allclose(o["aF"], aF, atol=1e-11)
            C_ref = acl + (2 / dt) * vF[:3] + (P - cp) / dt ** 2     # point_contact.hxx:67-86
            assert np.allclose(o["C"], C_ref, rtol=1e-12, atol=1e-9)
            g = lambda key: (lambda qq, vv, aa: fb.contact(qq, vv, aa, i, dt, cp)[key])
            assert np.allclose(o["J"], fd_x(lambda x: g("vF")(q, x, a), v), atol=1e-7)
            assert np.allclose(o["v_dq"], fd_q(fb, lambda x: g("vF")(x, v, a), q), atol=2e-6)
            assert np.allclose(o["a_dq"], fd_q(fb, lambda x: g("aF")(x, v, a), q), atol=2e-5)
            assert np.allclose(o["a_dv"], fd_x(lambda x: g("aF")(q, x, a), v), atol=2e-6)
            assert np.allclose(o["J"], fd_x(lambda x: g("aF")(q, v, x), a), atol=1e-7)
            # world-frame position Jacobian R J_lin (point_contact.hxx:195-201)
            assert np.allclose(R @ o["J"][:3], fd_q(fb, lambda x: g("P")(x, v, a), q), atol=1e-7)
            # Baumgarte derivatives: the reference's assembly (point_contact.hxx:89-144), sign quirk included
            sk = fbm.skew
            dq_ref = (o["a_dq"][:3] + sk(vF[3:]) @ o["v_dq"][:3] + sk(vF[:3]) @ o["v_dq"][3:]
                      + (2 / dt) * o["v_dq"][:3] + (1 / dt ** 2) * R @ o["J"][:3])
            dv_ref = o["a_dv"][:3] + sk(vF[3:]) @ o["J"][:3] + sk(vF[:3]) @ o["J"][3:] + (2 / dt) * o["J"][:3]
            assert np.allclose(o["dCdq"], dq_ref, rtol=1e-12, atol=1e-8)
            assert np.allclose(o["dCdv"], dv_ref, rtol=1e-12, atol=1e-9)
            assert np.array_equal(o["dCda"], o["J"][:3])
            # with the exact sign the assembly is the true derivative of the residual
            dq_true = dq_ref - 2 * sk(vF[:3]) @ o["v_dq"][3:]
            dv_true = dv_ref - 2 * sk(vF[:3]) @ o["J"][3:]
            assert np.allclose(dq_true, fd_q(fb, lambda x: g("C")(x, v, a), q), rtol=1e-5, atol=2e-3)
            assert np.allclose(dv_true, fd_x(lambda x: g("C")(q, x, a), v), rtol=1e-5, atol=1e-4)


def test_mjtjinv_is_the_kkt_inverse(fb):
    # test/robot/robot_test.cpp:545-565
    rng = np.random.default_rng(6)
    q, v, a = rand_state(rng)
    _, _, _, M = fb.rnea(q, v, a, derivatives=True)
    for active in ([0, 1, 2, 3], [1, 2], [3], []):
        J = np.concatenate([fb.contact(q, v, a, i, 0.05, np.zeros(3))["dCda"] for i in active]) if active else np.zeros((0, 18))
        out, info = fb.mjtjinv(M, J)
        assert info == 0
        n = 18 + J.shape[0]
        K = np.zeros((n, n))
        K[:18, :18] = M
        K[:18, 18:] = J.T
        K[18:, :18] = J
        assert np.allclose(out @ K, np.eye(n), atol=1e-9)
        assert np.allclose(out, out.T, atol=1e-12)

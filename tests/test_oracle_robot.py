"""Oracle vs independent formulations of the rigid-body dynamics (the pinocchio boundary).

The reference pins nothing here (its robot_test.cpp:414-531 compares against pinocchio itself), so
the oracle is validated against (i) a body-frame RNEA written independently in numpy, (ii) central
finite differences, (iii) the dense 6x6 world-frame derivative algorithm, (iv) structural
identities (dtau/da = M symmetric positive definite, hand-computable gravity torque)."""
import numpy as np
import pytest


def _rand(rng):
    return rng.uniform(-2.5, 2.5, 7), rng.uniform(-4, 4, 7), rng.uniform(-8, 8, 7)


def test_model_table_matches_urdf_facts(mirror):
    m = mirror.load_model()
    assert m["names"] == ["iiwa_joint_%d" % i for i in range(1, 8)]
    assert m["parent"] == [-1, 0, 1, 2, 3, 4, 5]
    # frame ids quoted by the reference: examples/iiwa14/task_space_ocp.cpp:67 (22), SURVEY.md (10)
    assert m["frames"][22] == "iiwa_link_ee_kuka" and m["frames"][10] == "iiwa_link_3"
    assert np.allclose(m["mass"], [4, 4, 3, 2.7, 1.7, 1.8, 0.3])
    assert np.allclose(m["q_max"], [2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239,
                                    2.96705972839, 2.09439510239, 3.05432619099])
    assert np.allclose(m["effort"], 300) and np.allclose(m["v_max"], 10)
    for R in m["R"]:
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-14)


def test_rnea_matches_body_frame_numpy(oracle, mirror):
    m = mirror.load_model()
    rng = np.random.default_rng(11)
    for _ in range(20):
        q, v, a = _rand(rng)
        assert np.allclose(oracle.rnea(q, v, a), mirror.rnea_body(m, q, v, a), rtol=1e-12, atol=1e-12)


def test_rnea_gravity_at_rest(oracle):
    # straight-up arm: gravity torque only from the small lateral com offsets
    tau = oracle.rnea(np.zeros(7), np.zeros(7), np.zeros(7))
    assert abs(tau[0]) < 1e-12 and abs(tau[6]) < 1e-12
    assert np.max(np.abs(tau)) < 0.05
    # horizontal arm (joint 2 at 90 deg): large torque on joint 2
    q = np.zeros(7); q[1] = np.pi / 2
    tau = oracle.rnea(q, np.zeros(7), np.zeros(7))
    assert abs(tau[1]) > 30.0


def test_rnea_derivatives_match_world_frame_numpy(oracle, mirror):
    m = mirror.load_model()
    rng = np.random.default_rng(12)
    for _ in range(10):
        q, v, a = _rand(rng)
        dq, dv, da = oracle.rnea_derivatives(q, v, a)
        tau, dq2, dv2, M2 = mirror.rnea_derivatives_world(m, q, v, a)      # dense 6x6 formulation
        _, dq3, dv3, M3 = mirror.rnea_derivatives_compact(m, q, v, a)     # structured formulation
        for x, y in ((dq, dq2), (dv, dv2), (da, M2), (dq, dq3), (dv, dv3), (da, M3)):
            assert np.max(np.abs(x - y)) <= 1e-12 * max(1.0, np.max(np.abs(y)))
        assert np.allclose(tau, oracle.rnea(q, v, a), rtol=1e-12, atol=1e-12)


def test_rnea_derivatives_match_finite_differences(oracle, mirror):
    m = mirror.load_model()
    rng = np.random.default_rng(13)
    for _ in range(5):
        q, v, a = _rand(rng)
        dq, dv, da = oracle.rnea_derivatives(q, v, a)
        fq, fv, fa = mirror.rnea_derivatives_fd(m, q, v, a)
        assert np.max(np.abs(dq - fq)) < 2e-6 * max(1.0, np.max(np.abs(dq)))
        assert np.max(np.abs(dv - fv)) < 2e-6 * max(1.0, np.max(np.abs(dv)))
        assert np.max(np.abs(da - fa)) < 2e-6


def test_mass_matrix_identities(oracle):
    # robot.hxx:496-499 mirrors the upper triangle: exactly symmetric; M is positive definite and
    # independent of v, a; rnea is affine in a with slope M
    rng = np.random.default_rng(14)
    q, v, a = _rand(rng)
    _, _, M = oracle.rnea_derivatives(q, v, a)
    assert np.array_equal(M, M.T)
    assert np.all(np.linalg.eigvalsh(M) > 0)
    _, _, M2 = oracle.rnea_derivatives(q, rng.uniform(-1, 1, 7), rng.uniform(-1, 1, 7))
    assert np.allclose(M, M2, rtol=1e-13, atol=1e-14)
    a2 = rng.uniform(-3, 3, 7)
    assert np.allclose(oracle.rnea(q, v, a2) - oracle.rnea(q, v, a), M @ (a2 - a), rtol=1e-11, atol=1e-11)


def test_splitmix_matches_bench_generator(oracle):
    import bench
    idx = np.array([0, 1, 2, 12345, 2 ** 40 + 7], dtype=np.uint64)
    ref = np.array([oracle.splitmix_uniform(bench.SEED, int(i)) for i in idx])
    assert np.array_equal(bench.splitmix_uniform(bench.SEED, idx), ref)
    assert np.all((ref >= 0) & (ref < 1))

"""Oracle restatement of idocp's UnParNMPCSolver against the identities the reference's own unit
tests assert (test/unocp/split_unkkt_matrix_inverter_test.cpp, split_unbackward_correction_test.cpp,
terminal_unparnmpc_test.cpp) and against an independent numpy re-derivation of one iteration."""
import numpy as np
import pytest

NV = 7


def _kkt_matrix(Q, dt):
    """[[0, F],[F^T, Q]] with F = [[0,-I,dt I],[dt I,0,-I]] (split_unkkt_matrix_inverter_test.cpp:42-50)."""
    eye, Z = np.eye(NV), np.zeros((NV, NV))
    F = np.block([[Z, -eye, dt * eye], [dt * eye, Z, -eye]])
    return np.block([[np.zeros((2 * NV, 2 * NV)), F], [F.T, Q]])


def test_kkt_matrix_inverter_identity(oracle):
    rng = np.random.default_rng(0)
    for dt in (0.05, 0.3):
        A = rng.standard_normal((21, 21))
        Q = A @ A.T + 21 * np.eye(21)
        info, Kinv = oracle.invert_unkkt(dt, Q)
        assert info == 0
        KKT = _kkt_matrix(Q, dt)
        assert np.abs(KKT @ Kinv - np.eye(35)).max() < 1e-10
        assert np.abs(Kinv - np.linalg.inv(KKT)).max() < 1e-10 * np.abs(Kinv).max()
        # LLT reads the lower triangle only: garbage above the diagonal changes nothing
        Q2 = Q.copy()
        Q2[np.triu_indices(21, 1)] = 1e30
        info2, Kinv2 = oracle.invert_unkkt(dt, Q2)
        assert info2 == 0 and np.array_equal(Kinv, Kinv2)


def test_kkt_matrix_inverter_reports_indefinite_q(oracle):
    Q = np.eye(21)
    Q[5, 5] = -1.0
    info, _ = oracle.invert_unkkt(0.05, Q)
    assert info == 6


def _solver(O, p, q, v):
    s = O.UnParNMPCSolver(p)
    s.set_solution("q", q)
    s.set_solution("v", v)
    s.init_backward_correction(0.0)
    return s


def test_single_stage_horizon_is_exact_newton(oracle):
    """N = 1: only the terminal stage, no corrections: the iteration is Newton's method on one KKT system."""
    O = oracle
    p = O.config_space_problem()
    p.N, p.T = 1, 0.05
    q = np.array([np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2])
    v = np.zeros(7)
    s = _solver(O, p, q, v)
    errs = []
    for _ in range(25):
        s.update_solution(0.0, q, v)
        s.compute_kkt_residual(0.0, q, v)
        errs.append(s.kkt_error())
    assert errs[-1] < 1e-8, errs


@pytest.mark.parametrize("N", [5, 20])
def test_converges_on_config_space_problem(oracle, N):
    O = oracle
    p = O.config_space_problem()
    p.N, p.T = N, 0.05 * N
    q = np.array([np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2])
    v = np.zeros(7)
    s = _solver(O, p, q, v)
    s.compute_kkt_residual(0.0, q, v)
    e0 = s.kkt_error()
    for _ in range(80):
        s.update_solution(0.0, q, v)
    s.compute_kkt_residual(0.0, q, v)
    assert s.kkt_error() < 1e-6 * e0
    assert s.chol_info() == 0


def test_coarse_update_and_corrections_match_numpy(oracle):
    """One updateSolution re-derived with numpy from the oracle's own stage linearisation:
    d = KKT^-1 r per stage, then the four correction sweeps of UnBackwardCorrection::backwardCorrection
    (src/unocp/unbackward_correction.cpp:100-134) written with dense blocks of the 35x35 inverse."""
    O = oracle
    p = O.benchmark_problem()
    p.N, p.T = 6, 0.3
    rng = np.random.default_rng(5)
    q, v = rng.uniform(-1, 1, 7), rng.uniform(-0.3, 0.3, 7)
    s = _solver(O, p, q, v)
    for _ in range(2):
        s.update_solution(0.0, q, v)
    N = p.N
    fields = ("lmd", "gmm", "a", "q", "v")
    before = {n: s.get_solution(n) for n in fields + ("u", "beta")}
    s.update_solution(0.0, q, v)
    Kinv = [s.get_kkt_inverse(i) for i in range(N)]
    for K in Kinv:   # symmetric up to rounding, and an actual inverse of a KKT matrix of that structure
        assert np.abs(K - K.T).max() <= 1e-9 * np.abs(K).max()
        assert np.abs(K[:14, :14] + np.linalg.inv(np.linalg.inv(-K[:14, :14]))).max() < 1e-6 * np.abs(K[:14, :14]).max()
    d = {n: s.get_direction("d" + n) for n in fields}
    primal = s.step_sizes()[0]
    after = {n: s.get_solution(n) for n in fields}
    for n in fields:
        assert np.allclose(after[n], before[n] + primal * d[n], rtol=0, atol=1e-12 * max(1.0, np.abs(after[n]).max()))
    # stage 0 has no forward correction, so its total direction is
    #   d_0 = -(KKT_0^-1 r_0) - KKT_0^-1[:, 21:35] (s_new_1 - s_1)(lmd, gmm)
    # recover r_0 = [Fq, Fv, la, lq, lv] from it and compare its state-equation part with the
    # backward-Euler defects of stage 0 (state_equation.hxx:224-236) computed directly
    xr_back0 = np.concatenate([d["lmd"][1], d["gmm"][1]])
    r0 = np.linalg.solve(Kinv[0], -(np.concatenate([d["lmd"][0], d["gmm"][0], d["a"][0], d["q"][0], d["v"][0]])
                                     + Kinv[0][:, 21:35] @ xr_back0))
    # r0 = [Fq, Fv, la, lq, lv]: its first 14 entries are the backward-Euler defects of stage 0
    dt = p.T / p.N
    Fq = q - before["q"][0] + dt * before["v"][0]
    Fv = v - before["v"][0] + dt * before["a"][0]
    assert np.allclose(r0[:7], Fq, atol=1e-7 * max(1.0, np.abs(Fq).max()))
    assert np.allclose(r0[7:14], Fv, atol=1e-7 * max(1.0, np.abs(Fv).max()))


@pytest.mark.parametrize("key,iters", [("unparnmpc_benchmark_reference_instance", 20), ("config_space_unparnmpc", 60)])
def test_golden_histories(oracle, key, iters):
    """examples/iiwa14/unparnmpc_benchmark.cpp instance (q = 2, v = 0) and the config-space problem through the
    ParNMPC driver: KKT error / step sizes of every iteration equal the committed golden vectors."""
    import json
    import os
    from conftest import GOLDEN
    O = oracle
    with open(os.path.join(GOLDEN, "solvers_golden.json")) as f:
        rec = json.load(f)[key]
    if key.startswith("unparnmpc_benchmark"):
        p = O.benchmark_problem()
    else:
        p = O.config_space_problem()
        p.N, p.T = 20, 1.0
    q0, v0 = np.array(rec["q0"]), np.array(rec["v0"])
    s = _solver(O, p, q0, v0)
    s.compute_kkt_residual(0.0, q0, v0)
    kkt = [s.kkt_error()]
    for it in range(iters):
        s.update_solution(0.0, q0, v0)
        st = s.step_sizes()
        assert st[0] == rec["primal"][it] and st[1] == rec["dual"][it]
        s.compute_kkt_residual(0.0, q0, v0)
        kkt.append(s.kkt_error())
    assert kkt == rec["kkt"]

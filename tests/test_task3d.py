"""TaskSpace3DCost / TimeVaryingTaskSpace3DCost: the oracle restatement against finite differences and the reference's own
formula, the CUDA sources (SIMT emulator here, the B200 under -m gpu) against the oracle."""
import numpy as np
import pytest

import task3d_scenarios as sc


def test_oracle_task3d_matches_the_reference_formula(oracle):
    """diff_3d = framePosition - q_3d_ref; J_3d = frameRotation * getFrameJacobian(LOCAL).topRows<3>() (task_space_3d_cost.cpp:
    92-97) from the oracle's independently validated frame kinematics; J_3d is the derivative of the frame position."""
    rng = np.random.default_rng(1)
    for _ in range(5):
        q = rng.uniform(-2, 2, 7)
        ref = sc.moving_target(rng.uniform(0, 2))
        diff, JJ = oracle.task_evaluate_kind(q, ref, 2)
        R, p, J = oracle.frame_kinematics(q)
        assert np.allclose(diff[:3], p - ref[9:], rtol=0, atol=1e-15) and np.all(diff[3:] == 0)
        assert np.allclose(JJ[:3], R @ J[:3], rtol=1e-13, atol=1e-15) and np.all(JJ[3:] == 0)
        eps = 1e-6
        for j in range(7):
            dq = np.zeros(7)
            dq[j] = eps
            pp = oracle.frame_kinematics(q + dq)[1]
            pm = oracle.frame_kinematics(q - dq)[1]
            assert np.allclose((pp - pm) / (2 * eps), JJ[:3, j], rtol=1e-7, atol=1e-9)


def test_oracle_task3d_converges(oracle):
    """Reaching a fixed position with the UnOCPSolver restatement: the KKT error falls by orders of magnitude and the
    end effector arrives."""
    prob = oracle.task_space_3d_problem()
    q0, v0 = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]), np.zeros(7)
    s = oracle.UnOCPSolver(prob)
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    s.set_task_ref(oracle.task_ref_table(sc.fixed_target, 0.0, prob.T, prob.N, "unocp"))
    s.compute_kkt_residual(0.0, q0, v0)
    k0 = s.kkt_error()
    for _ in range(40):
        s.update_solution(0.0, q0, v0, False)
    s.compute_kkt_residual(0.0, q0, v0)
    assert s.kkt_error() < 1e-6 * k0
    p_end = oracle.frame_kinematics(s.get_solution("q")[-1])[1]
    assert np.linalg.norm(p_end - sc.fixed_target(0)[9:]) < 2e-2


def test_task3d_emulator_matches_oracle(emu_lib, oracle):
    sc.run_task3d(emu_lib, oracle, batch=3, iters=2, N=6, T=0.3)


@pytest.mark.gpu
def test_task3d_gpu_matches_oracle(gpu_lib, oracle):
    first, last = sc.run_task3d(gpu_lib, oracle, batch=9, iters=6, N=30, T=1.5)
    assert np.all(np.isfinite(last))

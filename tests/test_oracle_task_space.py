"""Oracle restatement of TimeVaryingTaskSpace6DCost (src/cost/time_varying_task_space_6d_cost.cpp) and of the
pinocchio pieces behind it (frame placement, LOCAL frame Jacobian, log6, Jlog6) against independent numpy /
scipy evaluations, finite differences, and the committed golden vectors."""
import json
import os

import numpy as np
import pytest
from scipy.linalg import expm, logm

from conftest import GOLDEN


def _model():
    with open(os.path.join(GOLDEN, "model_iiwa14.json")) as f:
        return json.load(f)


def _fk_numpy(M, q):
    T = np.eye(4)
    for i in range(7):
        Tl = np.eye(4)
        Tl[:3, :3] = np.array(M["R"][i])
        Tl[:3, 3] = M["p"][i]
        c, s = np.cos(q[i]), np.sin(q[i])
        Rz = np.eye(4)
        Rz[:2, :2] = [[c, -s], [s, c]]
        T = T @ Tl @ Rz
    E = np.eye(4)
    E[:3, :3] = np.array(M["ee_R"])
    E[:3, 3] = M["ee_p"]
    return T @ E


def _hat6(x):
    v, w = x[:3], x[3:]
    H = np.zeros((4, 4))
    H[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    H[:3, 3] = v
    return H


def _vee6(H):
    return np.array([H[0, 3], H[1, 3], H[2, 3], H[2, 1], H[0, 2], H[1, 0]])


def test_canonical_acos(oracle):
    xs = np.concatenate([np.linspace(-1, 1, 4001), [0.5, -0.5, 0.4999999, 1 - 1e-12, -1 + 1e-12, 1.0, -1.0]])
    err = max(abs(oracle.canon_acos(x) - np.arccos(x)) for x in xs)
    assert err <= 4.5e-16
    assert oracle.canon_acos(1.5) == 0.0 and oracle.canon_acos(-1.5) == np.pi


def test_frame_placement_and_local_jacobian(oracle):
    M = _model()
    assert M["frames"][M["ee_frame"]] == "iiwa_link_ee_kuka" and M["ee_frame"] == 22   # task_space_ocp.cpp:67
    rng = np.random.default_rng(0)
    for _ in range(5):
        q = rng.uniform(-2, 2, 7)
        R, p, J = oracle.frame_kinematics(q)
        T = _fk_numpy(M, q)
        assert np.allclose(R, T[:3, :3], atol=1e-13) and np.allclose(p, T[:3, 3], atol=1e-13)
        eps = 1e-6
        for c in range(7):
            dq = np.zeros(7)
            dq[c] = eps
            col = _vee6(np.real(logm(np.linalg.inv(_fk_numpy(M, q - dq)) @ _fk_numpy(M, q + dq)))) / (2 * eps)
            assert np.allclose(J[:, c], col, atol=1e-7)


def test_log6_and_jlog6(oracle):
    M = _model()
    rng = np.random.default_rng(1)
    for _ in range(5):
        q = rng.uniform(-2, 2, 7)
        Rref = expm(_hat6(np.concatenate([np.zeros(3), rng.uniform(-1, 1, 3)])))[:3, :3]
        pref = rng.uniform(-0.5, 0.5, 3) + np.array([0.5, 0, 0.7])
        ref12 = np.concatenate([Rref.reshape(-1), pref])
        diff, JJ = oracle.task_evaluate(q, ref12)
        Tref = np.eye(4)
        Tref[:3, :3] = Rref
        Tref[:3, 3] = pref
        assert np.allclose(diff, _vee6(np.real(logm(np.linalg.inv(Tref) @ _fk_numpy(M, q)))), atol=1e-10)
        eps = 1e-6
        for c in range(7):
            dq = np.zeros(7)
            dq[c] = eps
            col = (oracle.task_evaluate(q + dq, ref12)[0] - oracle.task_evaluate(q - dq, ref12)[0]) / (2 * eps)
            assert np.allclose(JJ[:, c], col, atol=2e-7)


def test_log6_branches(oracle):
    """identity error (Taylor branch of log3 / log6 / Jlog6) and a rotation error next to pi (log3 special case)."""
    q = np.zeros(7)
    R, p, J = oracle.frame_kinematics(q)
    diff, JJ = oracle.task_evaluate(q, np.concatenate([R.reshape(-1), p]))
    assert np.array_equal(diff, np.zeros(6))
    assert np.allclose(JJ, J, atol=1e-15)          # Jlog6(identity) = I
    Rpi = R @ expm(_hat6(np.array([0, 0, 0, 0, 0, np.pi - 1e-3])))[:3, :3]
    diff, _ = oracle.task_evaluate(q, np.concatenate([Rpi.reshape(-1), p]))
    assert abs(np.linalg.norm(diff[3:]) - (np.pi - 1e-3)) < 1e-7


def test_weight_order_quirk(oracle):
    """set_q_6d_weight(position_weight, rotation_weight) stores head<3> = rotation_weight while diff_6d =
    [linear; angular] (time_varying_task_space_6d_cost.cpp:43-50, 73-75): the ROTATION weight multiplies the
    LINEAR part of the error.  With a pure translation error and only a position weight the cost gradient is zero."""
    O = oracle
    q = np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0.0])
    R, p, _ = O.frame_kinematics(q)
    ref = np.concatenate([R.reshape(-1), p + np.array([0.0, 0.05, 0.0])])   # pure translation error
    table = np.tile(ref, (3, 1))

    def kkt(pos_w, rot_w):
        pr = O.task_space_problem(N=2, T=0.1)
        for i in range(7):
            pr.v_weight[i] = pr.vf_weight[i] = pr.a_weight[i] = 0.0
        for k in range(3):
            pr.task_q_weight[k] = pr.task_qf_weight[k] = pos_w
            pr.task_q_weight[3 + k] = pr.task_qf_weight[3 + k] = rot_w
        s = O.UnOCPSolver(pr)
        s.set_solution("q", q)
        s.set_task_ref(table)
        s.compute_kkt_residual(0.0, q, np.zeros(7))
        return s.kkt_error()

    base = kkt(0.0, 0.0)
    assert kkt(1000.0, 0.0) == base          # position weight alone: no effect on a translation error
    assert kkt(0.0, 1000.0) > base + 1.0     # the rotation weight does act on it


def _golden():
    with open(os.path.join(GOLDEN, "solvers_golden.json")) as f:
        return json.load(f)


def test_task_space_kinematics_golden(oracle):
    for rec in _golden()["task_space_kinematics"]:
        diff, JJ = oracle.task_evaluate(rec["q"], rec["ref"])
        R, p, J = oracle.frame_kinematics(rec["q"])
        assert np.array_equal(diff, rec["diff"]) and np.array_equal(JJ, rec["JJ"])
        assert np.array_equal(R, rec["R"]) and np.array_equal(p, rec["p"]) and np.array_equal(J, rec["J"])


@pytest.mark.parametrize("key,kind", [("task_space_ocp_unocp", "unocp"), ("task_space_ocp_unparnmpc", "unparnmpc")])
def test_task_space_ocp_converges_and_matches_golden(oracle, key, kind):
    """examples/iiwa14/task_space_ocp.cpp (T = 6, N = 120, 30 iterations): the KKT error drops by many orders and
    the end effector follows the circle; the history equals the committed golden vector."""
    O = oracle
    rec = _golden()[key]
    p = O.task_space_problem()
    cls = O.UnOCPSolver if kind == "unocp" else O.UnParNMPCSolver
    s = cls(p)
    q0, v0 = np.array(rec["q0"]), np.array(rec["v0"])
    s.set_solution("q", q0)
    s.set_solution("v", v0)
    s.set_task_ref(O.task_ref_table(O.task_space_ref, 0.0, p.T, p.N, kind))
    if kind != "unocp":
        s.init_backward_correction(0.0)
    s.compute_kkt_residual(0.0, q0, v0)
    kkt = [s.kkt_error()]
    for it in range(30):
        s.update_solution(0.0, q0, v0)
        st = s.step_sizes()
        assert st[0] == rec["primal"][it] and st[1] == rec["dual"][it]
        s.compute_kkt_residual(0.0, q0, v0)
        kkt.append(s.kkt_error())
    assert kkt == rec["kkt"]
    assert kkt[-1] < (1e-4 if kind == "unocp" else 5e-2) * kkt[0]
    if kind == "unocp":
        qs = s.get_solution("q")
        for i in (40, 80, 120):
            _, pe, _ = O.frame_kinematics(qs[i])
            assert np.linalg.norm(pe - O.task_space_ref(0.05 * i)[9:]) < 5e-3

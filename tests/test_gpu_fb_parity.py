"""ANYmal OCPSolver on the GPU (through the C-ABI) vs the oracle: bit-exact per iteration on a batch of perturbed
initial states, identical KKT-error histories to convergence, and size-independent properties at a larger batch."""
import numpy as np
import pytest

import anymal_problems as ap
from test_emu_fb_parity import DIR, EXP, KKT, RIC, SOL, compare, contact_distance_scenario, nonlinear_cone_scenario

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


def perturbed_states(fb, pr, batch, seed):
    rng = np.random.default_rng(seed)
    q0 = np.zeros((batch, 19))
    v0 = np.zeros((batch, 18))
    for b in range(batch):
        dq = np.concatenate([rng.uniform(-0.01, 0.01, 3), rng.uniform(-0.02, 0.02, 3), rng.uniform(-0.02, 0.02, 12)])
        q0[b] = fb.integrate(pr.q0, dq)
        v0[b] = rng.uniform(-0.1, 0.1, 18)
    q0[0], v0[0] = pr.q0, pr.v0
    return q0, v0


def test_batch_iterations_bit_exact(fb, gpu_lib):
    pr = ap.TrottingProblem()
    B = 6
    q0, v0 = perturbed_states(fb, pr, B, 1)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    for it in range(4):
        solver.computeKKTResidual(0.0, q0, v0)
        kkt = solver.KKTError()
        for b, o in enumerate(oracles):
            o.compute_kkt_residual(0.0, q0[b], v0[b])
            assert kkt[b] == o.kkt_error(), (it, b, kkt[b], o.kkt_error())
        solver.updateSolution(0.0, q0, v0)
        steps = solver.stepSizes()
        for b, o in enumerate(oracles):
            assert o.update_solution(0.0, q0[b], v0[b]) == 0
            assert np.array_equal(steps[b], o.step_sizes())
            for names in (KKT + EXP, RIC, DIR, SOL):
                assert compare(o, solver, fb, names, b=b) == [], (it, b)


def test_nonlinear_cones_and_acceleration_limits_bit_exact(fb, gpu_lib):
    # FrictionCone, ImpulseFrictionCone, JointAcceleration{Lower,Upper}Limit on a batch of perturbed states
    pr = ap.TrottingProblem()
    nonlinear_cone_scenario(fb, gpu_lib, perturbed_states(fb, pr, 5, 11))


@pytest.mark.parametrize("mode", [1, 2])
def test_contact_distance_bit_exact(fb, gpu_lib, mode):
    # ContactDistance, literal (1) and consistent (2), on a batch of perturbed states
    pr = ap.TrottingProblem()
    contact_distance_scenario(fb, gpu_lib, mode, perturbed_states(fb, pr, 5, 13))


def test_all_f3_components_together_bit_exact(fb, gpu_lib):
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem())
    contact_distance_scenario(fb, gpu_lib, 2, perturbed_states(fb, pr, 4, 17), iters=3, problem=pr)


def test_convergence_history_identical(fb, gpu_lib):
    # examples/anymal/anymal_trotting.cpp: 25 iterations; iteration-by-iteration KKT errors equal the oracle's bits
    pr = ap.TrottingProblem()
    B = 4
    q0, v0 = perturbed_states(fb, pr, B, 2)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    for it in range(25):
        solver.computeKKTResidual(0.0, q0, v0)
        kkt = solver.KKTError()
        for b, o in enumerate(oracles):
            o.compute_kkt_residual(0.0, q0[b], v0[b])
            assert kkt[b] == o.kkt_error(), (it, b)
            o.update_solution(0.0, q0[b], v0[b])
        solver.updateSolution(0.0, q0, v0)
    solver.computeKKTResidual(0.0, q0, v0)
    assert np.all(solver.KKTError() < 1e-8)
    e_last = len(solver.chain()) - 1
    for b, o in enumerate(oracles):
        assert np.array_equal(solver.get(e_last, "q")[b], o.get(e_last, "q"))


def test_flight_phase_bit_exact(fb, gpu_lib):
    pr = ap.JumpingProblem(0.1, 0.6, 0.75, 1.3, 26)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=2)
    o = pr.make_oracle(fb)
    for it in range(5):
        o.update_solution(0.0, pr.q0, pr.v0)
        solver.updateSolution(0.0, pr.q0, pr.v0)
        for names in (KKT + EXP, RIC, DIR, SOL):
            assert compare(o, solver, fb, names, b=1) == [], it


def test_large_batch_properties(fb, gpu_lib):
    # size-independent properties at a batch the oracle would need minutes for: identical instances give identical
    # bits wherever they sit in the batch; the KKT error decreases to convergence for every instance
    pr = ap.TrottingProblem()
    B = 256
    q0, v0 = perturbed_states(fb, pr, 8, 3)
    q0 = np.tile(q0, (B // 8, 1))
    v0 = np.tile(v0, (B // 8, 1))
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    hist = []
    for it in range(25):
        solver.computeKKTResidual(0.0, q0, v0)
        hist.append(solver.KKTError())
        solver.updateSolution(0.0, q0, v0)
    hist = np.array(hist)
    assert np.all(np.isfinite(hist)) and np.all(hist[-1] < 1e-7)
    assert np.array_equal(hist[:, :8], hist[:, 8:16]) and np.array_equal(hist[:, :8], hist[:, -8:])
    Kq, Kv = solver.getStateFeedbackGain(3)
    assert Kq.shape == (B, 12, 18) and np.array_equal(Kq[:8], Kq[-8:]) and np.all(np.isfinite(Kv))
    e_mid = len(solver.chain()) // 2
    u = solver.get(e_mid, "u")
    assert np.array_equal(u[:8], u[-8:])
    assert solver.launchCount() > 0


def test_filter_line_search_bit_exact(fb, gpu_lib):
    # LineSearch for OCPSolver (line_search.hpp:62-93): per-instance step sizes, filters and iterates equal the oracle's
    pr = ap.JumpingProblem(0.1, 0.6, 0.75, 1.3, 26)
    B = 4
    q0, v0 = perturbed_states(fb, pr, B, 5)
    q0 = np.array([fb.integrate(pr.q0, 0.25 * fb.subtract(q0[b], pr.q0)) for b in range(B)])
    v0 = 0.25 * v0
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    seen = set()
    compared = 0
    for it in range(6):
        solver.updateSolution(0.0, q0, v0, True)
        steps = solver.stepSizes()
        rcs = [o.update_solution(0.0, q0[b], v0[b], True) for b, o in enumerate(oracles)]
        for b, o in enumerate(oracles):
            assert np.array_equal(steps[b], o.step_sizes()), (it, b, steps[b], o.step_sizes())
            seen.add(float(steps[b][0]))
        if any(rcs):    # the 0.05 floor of the search can leave the interior (upstream behaviour): stop comparing
            break
        for b, o in enumerate(oracles):
            assert compare(o, solver, fb, SOL, b=b) == [], (it, b)
        compared += 1
    assert compared >= 3 and len(seen) > 2       # instances backtracked differently
    solver.clearLineSearchFilter()
    for o in oracles:
        o.clear_line_search_filter()
    solver.updateSolution(0.0, q0, v0, True)
    for b, o in enumerate(oracles):
        o.update_solution(0.0, q0[b], v0[b], True)
        assert np.array_equal(solver.stepSizes()[b], o.step_sizes())


def test_running_gait_with_flight_phases_bit_exact(fb, gpu_lib):
    # examples/anymal/anymal_running.cpp shortened to one stride: 8 impulses, 5 lifts, flight phases (dimf = 0), the
    # switching constraint on lift stages, a time-varying configuration reference
    pr = ap.RunningProblem(steps=1)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=2)
    o = pr.make_oracle(fb)
    o.set_threads(8)
    kinds = [c["kind"] for c in solver.chain()]
    assert kinds.count(fb.K_IMPULSE) == 8 and kinds.count(fb.K_LIFT) == 5
    assert [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in solver.chain()] == [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in o.chain()]
    for it in range(3):
        o.compute_kkt_residual(0.0, pr.q0, pr.v0)
        solver.computeKKTResidual(0.0, pr.q0, pr.v0)
        assert solver.KKTError()[1] == o.kkt_error()
        assert o.update_solution(0.0, pr.q0, pr.v0) == 0
        solver.updateSolution(0.0, pr.q0, pr.v0)
        for names in (KKT + EXP, RIC, DIR, SOL):
            assert compare(o, solver, fb, names, b=1) == [], it

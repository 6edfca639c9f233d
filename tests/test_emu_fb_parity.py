"""The ANYmal OCPSolver kernels (idocp_b200/csrc/fb_kernels.cuh) compiled with g++ against the SIMT emulator vs the
oracle (oracle/fb_ocp.c): bit-exact on every quantity of every stage.  CPU-side check of the cooperative code; the
same comparison runs on the real GPU in tests/test_gpu_fb_parity.py."""
import numpy as np
import pytest

import anymal_problems as ap

SOL = ["q", "v", "a", "u", "f", "lmd", "gmm", "beta", "mu", "nu_passive", "xi"]
DIR = ["dq", "dv", "du", "daf", "dbetamu", "dlmd", "dgmm", "dnu_passive", "dxi"]
# condensed KKT record; Qxx / Qxu / Quu / lu are left out: the oracle's Riccati sweep updates them in place
# (backward_riccati_recursion_factorizer.hxx:44-111) while the kernels keep the linearisation record untouched --
# they are covered through K, k, P, s, which are functions of exactly those blocks
KKT = ["lq", "lv", "lu_passive", "Fq", "Fv", "Fvq", "Fvv", "Fvu", "Fqq6", "Fqv6", "Fqq_prev_inv"]
EXP = ["MJtJinv", "MJ_dIDC", "MJ_IDC", "Qafqv", "Qafu"]
RIC = ["K", "k", "Pqq", "Pqv", "Pvv", "sq", "sv"]


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


def masked(name, arr, c, fb):
    """Entries the path defines (stacked blocks are only meaningful up to dimf / dimi)."""
    arr = np.array(arr, dtype=float)
    nvf, dimf, dimi = 18 + c["dimf"], c["dimf"], c["dimi"]
    if name in ("daf", "dbetamu", "MJ_IDC"):
        arr[nvf:] = 0
    elif name == "MJtJinv":
        m = arr.reshape(30, 30); m[nvf:] = 0; m[:, nvf:] = 0
    elif name in ("MJ_dIDC", "Qafqv", "Qafu"):
        arr.reshape(30, -1)[nvf:] = 0
    elif name in ("dxi", "xi"):
        arr[dimi:] = 0
    return arr


def compare(ocp, solver, fb, names, b=0, only_kinds=None):
    bad = []
    for e, c in enumerate(ocp.chain()):
        if only_kinds is not None and c["kind"] not in only_kinds:
            continue
        for nm in names:
            if c["kind"] == fb.K_TERMINAL and nm not in ("q", "v", "lmd", "gmm", "dq", "dv", "dlmd", "dgmm", "lq", "lv", "Pqq", "Pvv",
                                                         "sq", "sv", "Fqq_prev_inv"):
                continue
            if c["kind"] == fb.K_IMPULSE and nm in ("u", "du", "nu_passive", "dnu_passive", "lu", "lu_passive", "Qxu", "Quu", "Fvu",
                                                    "Fqv6", "Qafu", "K", "k", "xi", "dxi"):
                continue
            if nm in ("xi", "dxi") and c["dimi"] == 0:
                continue
            x = masked(nm, ocp.get(e, nm), c, fb)
            y = masked(nm, solver.get(e, nm)[b], c, fb)
            if nm == "Qxx":    # the Qvq block is dead storage on both sides (the recursion rebuilds it from Qqv)
                x.reshape(36, 36)[18:, :18] = 0
                y.reshape(36, 36)[18:, :18] = 0
            if not np.array_equal(x, y):
                bad.append((e, c["kind"], nm, float(np.nanmax(np.abs(x - y)))))
    return bad


def test_trotting_iterations_bit_exact(fb, emu_lib):
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    solver = ap.make_product_solver(pr, emu_lib, fb, batch=1)
    ch = solver.chain()
    assert [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ch] == [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ocp.chain()]
    assert np.array_equal([c["dt"] for c in ch], [c["dt"] for c in ocp.chain()])
    assert compare(ocp, solver, fb, SOL) == []
    for e in range(len(ch)):
        assert np.array_equal(ocp.get(e, "slack"), solver.get(e, "slack")[0]), e
        assert np.array_equal(ocp.get(e, "dual"), solver.get(e, "dual")[0]), e
    for it in range(3):
        ocp.compute_kkt_residual(0.0, pr.q0, pr.v0)
        solver.computeKKTResidual(0.0, pr.q0, pr.v0)
        assert solver.KKTError()[0] == ocp.kkt_error(), (it, solver.KKTError()[0], ocp.kkt_error())
        assert ocp.update_solution(0.0, pr.q0, pr.v0) == 0
        solver.updateSolution(0.0, pr.q0, pr.v0)
        assert compare(ocp, solver, fb, KKT + EXP) == [], it
        assert compare(ocp, solver, fb, RIC) == [], it
        assert compare(ocp, solver, fb, DIR) == [], it
        assert np.array_equal(solver.stepSizes()[0], ocp.step_sizes())
        assert compare(ocp, solver, fb, SOL) == [], it


def nonlinear_cone_scenario(fb, lib, batch_states=None):
    """FrictionCone + ImpulseFrictionCone + JointAcceleration{Lower,Upper}Limit through the kernels vs the
    oracle: every field, slack / dual of all 140 rows, KKT errors, step sizes; the first two iterations with the filter line search."""
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem())
    B = 1 if batch_states is None else len(batch_states[0])
    q0 = np.tile(pr.q0, (B, 1)) if batch_states is None else batch_states[0]
    v0 = np.tile(pr.v0, (B, 1)) if batch_states is None else batch_states[1]
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    solver = ap.make_product_solver(pr, lib, fb, batch=B, q0=q0, v0=v0)
    ne = len(solver.chain())

    def check_slack_dual(tag):
        for e in range(ne):
            for nm in ("slack", "dual"):
                got = np.asarray(solver.get(e, nm))
                assert got.shape[1] == 140
                for b, o in enumerate(oracles):
                    assert np.array_equal(o.get(e, nm), got[b]), (tag, e, nm, b)
    check_slack_dual("init")
    for it in range(4):
        ls = it < 2      # (with the search later the 0.05 floor exceeds the fraction-to-boundary step here and LLT(G) fails)
        solver.computeKKTResidual(0.0, q0, v0)
        kkt = solver.KKTError()
        for b, o in enumerate(oracles):
            o.compute_kkt_residual(0.0, q0[b], v0[b])
            assert kkt[b] == o.kkt_error(), (it, b, kkt[b], o.kkt_error())
        solver.updateSolution(0.0, q0, v0, ls)
        steps = solver.stepSizes()
        for b, o in enumerate(oracles):
            assert o.update_solution(0.0, q0[b], v0[b], ls) == 0
            assert np.array_equal(steps[b], o.step_sizes()), (it, b, steps[b], o.step_sizes())
            for names in (KKT + EXP, RIC, DIR, SOL):
                assert compare(o, solver, fb, names, b=b) == [], (it, b)
        check_slack_dual(it)
    # the acceleration limit and the cone rows are live: slack of the first grid stage
    sl = np.asarray(solver.get(0, "slack"))
    assert np.all(sl[:, 72:80] > 0) and np.all(sl[:, 80:92] == 0) and np.all(sl[:, 112:136] > 0)


def test_nonlinear_cones_and_acceleration_limits_bit_exact(fb, emu_lib):
    nonlinear_cone_scenario(fb, emu_lib)


def contact_distance_scenario(fb, lib, mode, batch_states=None, iters=4, problem=None):
    """ContactDistance (src/constraints/contact_distance.cpp) through the kernels vs the oracle: mode 1 = the
    reference literally (row 2 of the LOCAL frame Jacobian), mode 2 = the consistent variant; every field, all 140 constraint rows,
    KKT errors, step sizes; mode 2 starts with two iterations of the filter line search."""
    pr = problem or ap.TrottingProblem()
    pr.problem.enable_distance = mode
    B = 1 if batch_states is None else len(batch_states[0])
    q0 = np.tile(pr.q0, (B, 1)) if batch_states is None else batch_states[0]
    v0 = np.tile(pr.v0, (B, 1)) if batch_states is None else batch_states[1]
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    solver = ap.make_product_solver(pr, lib, fb, batch=B, q0=q0, v0=v0)
    ne = len(solver.chain())

    def check_slack_dual(tag):
        for e in range(ne):
            for nm in ("slack", "dual"):
                got = np.asarray(solver.get(e, nm))
                assert got.shape[1] == 140
                for b, o in enumerate(oracles):
                    assert np.array_equal(o.get(e, nm), got[b]), (tag, e, nm, b)
    check_slack_dual("init")
    sl = np.asarray(solver.get(5, "slack"))
    assert np.all(sl[:, 136:140] > 0)            # grid stage 5: position level, the four contact distances are live
    assert np.all(np.asarray(solver.get(0, "slack"))[:, 136:140] == 0)     # stage 0: not yet (time stage < 2)
    for it in range(iters):
        ls = mode == 2 and it < 2     # (mode 1 diverges: with the search its 0.05 floor soon breaks LLT(G))
        solver.computeKKTResidual(0.0, q0, v0)
        kkt = solver.KKTError()
        for b, o in enumerate(oracles):
            o.compute_kkt_residual(0.0, q0[b], v0[b])
            assert kkt[b] == o.kkt_error(), (it, b, kkt[b], o.kkt_error())
        solver.updateSolution(0.0, q0, v0, ls)
        steps = solver.stepSizes()
        for b, o in enumerate(oracles):
            assert o.update_solution(0.0, q0[b], v0[b], ls) == 0
            assert np.array_equal(steps[b], o.step_sizes()), (it, b, steps[b], o.step_sizes())
            for names in (KKT + EXP, RIC, DIR, SOL):
                assert compare(o, solver, fb, names, b=b) == [], (it, b)
        check_slack_dual(it)


@pytest.mark.parametrize("mode", [1, 2])
def test_contact_distance_bit_exact(fb, emu_lib, mode):
    contact_distance_scenario(fb, emu_lib, mode)


def test_all_f3_components_together_bit_exact(fb, emu_lib):
    # nonlinear cones + acceleration limits + contact distances on one problem: all eleven components, 140 rows
    contact_distance_scenario(fb, emu_lib, 2, problem=ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem()), iters=3)


def test_filter_line_search_bit_exact(fb, emu_lib):
    # LineSearch::computeStepSize for OCPSolver (line_search.hpp:62-93) on a problem where the search backtracks
    pr = ap.JumpingProblem(0.1, 0.6, 0.75, 1.3, 26)
    ocp = pr.make_oracle(fb)
    solver = ap.make_product_solver(pr, emu_lib, fb, batch=1)
    steps = []
    for it in range(4):
        assert ocp.update_solution(0.0, pr.q0, pr.v0, True) == 0
        solver.updateSolution(0.0, pr.q0, pr.v0, True)
        assert np.array_equal(solver.stepSizes()[0], ocp.step_sizes()), (it, solver.stepSizes()[0], ocp.step_sizes())
        steps.append(ocp.step_sizes()[0])
        assert compare(ocp, solver, fb, SOL) == [], it
    assert min(steps) < 0.9          # the filter rejected at least one full fraction-to-boundary step
    assert ocp.filter_size() >= 2
    # clearLineSearchFilter: the next call re-evaluates the current point first
    ocp.clear_line_search_filter()
    solver.clearLineSearchFilter()
    assert ocp.update_solution(0.0, pr.q0, pr.v0, True) == 0
    solver.updateSolution(0.0, pr.q0, pr.v0, True)
    assert np.array_equal(solver.stepSizes()[0], ocp.step_sizes())
    assert compare(ocp, solver, fb, SOL) == []


def test_receding_horizon_calls_bit_exact(fb, emu_lib):
    # the MPC caller one step outside the path: updateSolution at advancing t re-discretises the
    # schedule (stage times, the shortened stage before an event, slot <-> stage mapping), events leave the horizon
    # through popFrontContactStatus and enter through pushBackContactStatus (ocp_solver.cpp:174-194)
    pr = ap.TrottingProblem()
    ocp = pr.make_oracle(fb)
    solver = ap.make_product_solver(pr, emu_lib, fb, batch=1)
    cs = ocp.cs
    q, v = pr.q0.copy(), pr.v0.copy()
    for t in (0.0, 0.013, 0.04, 0.31, 0.47):
        pr.set_references(ocp, t)
        assert ocp.update_solution(t, q, v) == 0
        solver.updateSolution(t, q, v)
        assert [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in solver.chain()] == \
               [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ocp.chain()]
        assert np.array_equal([c["dt"] for c in solver.chain()], [c["dt"] for c in ocp.chain()])
        assert np.array_equal(solver.stepSizes()[0], ocp.step_sizes()), t
        assert compare(ocp, solver, fb, SOL) == [], t
        # the plant moves: the next initial state is the second stage of the current solution
        q, v = ocp.get(1, "q"), ocp.get(1, "v")
    # the first event (lift at 0.5) has passed: drop it, append a new touch-down at the end of the horizon
    cs.pop_front()
    solver.popFrontContactStatus()
    a, pts = cs.phase(cs.counts()[0] - 1)
    pts = pts.copy()
    pts[0, 0] += pr.step_length
    pts[3, 0] += pr.step_length
    t = 0.52
    assert cs.push_back([1, 0, 0, 1], t + pr.T - 0.02, pts) == 0
    solver.pushBackContactStatus([1, 0, 0, 1], t + pr.T - 0.02, pts)
    for t in (0.52, 0.55):
        pr.set_references(ocp, t)
        rc = ocp.update_solution(t, q, v)
        solver.updateSolution(t, q, v)
        assert [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in solver.chain()] == \
               [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ocp.chain()]
        assert np.array_equal(solver.stepSizes()[0], ocp.step_sizes()), t
        if rc == 0:
            assert compare(ocp, solver, fb, SOL) == [], t


def test_state_feedback_gain_matches_oracle_emu(fb, emu_lib):
    import fb_scenarios
    fb_scenarios.run_state_feedback_gain(emu_lib, fb, batch=2)


def test_receding_horizon_batch_emu(fb, emu_lib):
    import fb_scenarios
    fb_scenarios.run_receding_horizon(emu_lib, fb, batch=2)


def test_event_before_t_is_an_error_emu(fb, emu_lib):
    import fb_scenarios
    fb_scenarios.run_event_before_t_is_an_error(emu_lib, fb)


def test_event_entering_horizon_keeps_constraints_emu(fb, emu_lib):
    import fb_scenarios
    fb_scenarios.run_event_entering_horizon_keeps_constraints(emu_lib, fb)


def test_batched_mpc_ticks_emu(fb, emu_lib):
    import fb_scenarios
    fb_scenarios.run_mpc_ticks(emu_lib, fb, batch=2, ticks=(0.0, 0.3, 0.52), iterations=1)

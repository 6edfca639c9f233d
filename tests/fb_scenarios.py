"""Scenario bodies shared by the emulator tests (CPU, tiny batches) and the GPU tests (BASELINE-sized batches): the
library under test is a parameter, the oracle is always the checker."""
import os

import numpy as np

import anymal_problems as ap
from test_emu_fb_parity import DIR, SOL

THREADS = os.cpu_count() or 1
_TERMINAL_OK = ("q", "v", "lmd", "gmm", "dq", "dv", "dlmd", "dgmm", "lq", "lv", "Pqq", "Pvv", "sq", "sv", "Fqq_prev_inv")
_IMPULSE_SKIP = ("u", "du", "nu_passive", "dnu_passive", "lu", "lu_passive", "Qxu", "Quu", "Fvu", "Fqv6", "Qafu", "K", "k",
                 "xi", "dxi")
_MASKED = ("daf", "dbetamu", "MJ_IDC", "MJtJinv", "MJ_dIDC", "Qafqv", "Qafu", "dxi", "xi")


def masked_batch(name, arr, c):
    """masked() of test_emu_fb_parity.py for (batch, size) arrays: entries the path defines (stacked blocks are only
    meaningful up to dimf / dimi)."""
    arr = np.array(arr, dtype=float)
    nvf, dimi = 18 + c["dimf"], c["dimi"]
    if name in ("daf", "dbetamu", "MJ_IDC"):
        arr[:, nvf:] = 0
    elif name == "MJtJinv":
        m = arr.reshape(-1, 30, 30)
        m[:, nvf:] = 0
        m[:, :, nvf:] = 0
    elif name in ("MJ_dIDC", "Qafqv", "Qafu"):
        arr.reshape(arr.shape[0], 30, -1)[:, nvf:] = 0
    elif name in ("dxi", "xi"):
        arr[:, dimi:] = 0
    return arr


def compare_batch(oracles, solver, fb, names, rows=None):
    """Every field in `names` of every chain element, all instances at once (rows: boolean mask of the instances to
    compare).  Returns [(element, kind, field, number of differing instances)]."""
    bad = []
    for e, c in enumerate(solver.chain()):
        for nm in names:
            if c["kind"] == fb.K_TERMINAL and nm not in _TERMINAL_OK:
                continue
            if c["kind"] == fb.K_IMPULSE and nm in _IMPULSE_SKIP:
                continue
            if nm in ("xi", "dxi") and c["dimi"] == 0:
                continue
            x = fb.batch_get(oracles, e, nm)
            y = np.asarray(solver.get(e, nm), dtype=float).reshape(x.shape)
            if nm in _MASKED:
                x, y = masked_batch(nm, x, c), masked_batch(nm, y, c)
            if rows is not None:
                x, y = x[rows], y[rows]
            if not np.array_equal(x, y, equal_nan=True):
                same = (x == y) | (np.isnan(x) & np.isnan(y))
                bad.append((e, c["kind"], nm, int((~np.all(same, axis=1)).sum())))
    return bad


def anymal_states(pr, count, seed):
    from idocp_b200 import problems as P
    return P.anymal_initial_states(0, count, q_nominal=pr.q0, seed=seed)


def chain_signature(ch):
    return [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ch]


def run_state_feedback_gain(lib, fb, batch):
    """OCPSolver::getStateFeedbackGain (riccati_recursion_solver.cpp:254-260): K of the LQR policy of a time stage."""
    pr = ap.TrottingProblem()
    q0, v0 = anymal_states(pr, batch, 20240004)
    solver = ap.make_product_solver(pr, lib, fb, batch=batch, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(batch)]
    for _ in range(2):
        solver.updateSolution(0.0, q0, v0)
        fb.batch_update_solution(oracles, 0.0, q0, v0, False, THREADS)
    ch = solver.chain()
    for stage in (0, 3, 11, pr.N - 1):
        e = [k for k, c in enumerate(ch) if c["kind"] == fb.K_GRID and c["index"] == stage][0]
        Kq, Kv = solver.getStateFeedbackGain(stage)
        K = fb.batch_get(oracles, e, "K").reshape(batch, 12, 36)
        assert np.array_equal(Kq, K[:, :, :18]) and np.array_equal(Kv, K[:, :, 18:]), stage


def run_receding_horizon(lib, fb, batch):
    """MPC-style use (ocp_solver.cpp:174-194): updateSolution at advancing t re-discretises the schedule, the plant moves
    (next x0 = second stage of the current solution, per instance), the first phase is popped once its event has
    passed and a new touch-down is pushed at the end of the horizon; 7 ticks, every instance bit for bit."""
    pr = ap.TrottingProblem()
    q, v = anymal_states(pr, batch, 20240004)
    q[0], v[0] = pr.q0, pr.v0
    solver = ap.make_product_solver(pr, lib, fb, batch=batch, q0=q, v0=v)
    oracles = [pr.make_oracle(fb, q0=q[b], v0=v[b]) for b in range(batch)]

    def tick(t, q, v):
        for o in oracles:
            pr.set_references(o, t)
        rcs = [o.update_solution(t, q[b], v[b]) for b, o in enumerate(oracles)]
        solver.updateSolution(t, q, v)
        assert chain_signature(solver.chain()) == chain_signature(oracles[0].chain()), t
        assert np.array_equal([c["dt"] for c in solver.chain()], [c["dt"] for c in oracles[0].chain()]), t
        assert np.array_equal(solver.stepSizes(), np.array([o.step_sizes() for o in oracles]), equal_nan=True), t
        ok = np.array(rcs) == 0
        assert compare_batch(oracles, solver, fb, SOL + DIR, rows=ok) == [], t
        return fb.batch_get(oracles, 1, "q"), fb.batch_get(oracles, 1, "v"), ok

    ticks = 0
    for t in (0.0, 0.013, 0.04, 0.31, 0.47):
        q, v, ok = tick(t, q, v)
        assert ok.all(), t
        ticks += 1
    for o in oracles:
        o.cs.pop_front()
    solver.popFrontContactStatus()
    cs = oracles[0].cs
    a, pts = cs.phase(cs.counts()[0] - 1)
    pts = pts.copy()
    pts[0, 0] += pr.step_length
    pts[3, 0] += pr.step_length
    t = 0.52
    for o in oracles:
        assert o.cs.push_back([1, 0, 0, 1], t + pr.T - 0.02, pts) == 0
    solver.pushBackContactStatus([1, 0, 0, 1], t + pr.T - 0.02, pts)
    for t in (0.52, 0.55):
        q, v, ok = tick(t, q, v)
        ticks += 1
    assert ticks == 7


def run_event_before_t_is_an_error(lib, fb):
    """time advanced past an event without popFrontContactStatus -> the discretisation is ill-defined
    (the reference asserts, ocp_discretizer.hxx:62-72); the product must refuse instead of silently solving a chain
    that dropped every later event.  The opt-out reproduces the Release-build behaviour."""
    import idocp_b200 as I
    pr = ap.TrottingProblem()
    solver = ap.make_product_solver(pr, lib, fb, batch=1)
    o = pr.make_oracle(fb)
    t = pr.t_start + 0.07          # the first event (t_start) now lies before t
    assert o.discretize(t) == -3
    try:
        solver.updateSolution(t, pr.q0, pr.v0)
    except I.Idocp_b200Error as e:
        assert "popFrontContactStatus" in str(e)
    else:
        raise AssertionError("ill-defined discretisation was accepted")
    try:
        solver.computeKKTResidual(t, pr.q0, pr.v0)
    except I.Idocp_b200Error:
        pass
    else:
        raise AssertionError("ill-defined discretisation was accepted by computeKKTResidual")
    solver.setStrictDiscretization(False)
    solver.updateSolution(t, pr.q0, pr.v0)          # runs on, like the reference built with NDEBUG
    solver.setStrictDiscretization(True)
    solver.popFrontContactStatus()
    solver.updateSolution(t, pr.q0, pr.v0)          # the documented fix


def run_event_entering_horizon_keeps_constraints(lib, fb, batch=1):
    """an event scheduled beyond the horizon at initConstraints time enters it later (receding horizon).
    Its impulse / aux / lift stages must carry active, initialised constraints (OCPLinearizer::initConstraints covers
    contact_sequence.numImpulseEvents(), ocp_linearizer.cpp:40-68), not silently disabled ones."""
    pr = ap.TrottingProblem()
    pr.max_num_impulse = pr.problem.max_num_impulse = 4
    cs_extra_t = pr.T + 0.3        # a fourth event, 0.3 s beyond the end of the horizon at t = 0
    q, v = anymal_states(pr, batch, 20240004)

    base_sequence = pr.contact_sequence

    def sequence_with_extra(fbm):
        cs = base_sequence(fbm)
        a, pts = cs.phase(cs.counts()[0] - 1)
        pts = pts.copy()
        pts[0, 0] += pr.step_length
        pts[3, 0] += pr.step_length
        assert cs.push_back([1, 0, 0, 1], cs_extra_t, pts) == 0     # a touch-down: impulse + aux stages
        return cs
    pr.contact_sequence = sequence_with_extra
    solver = ap.make_product_solver(pr, lib, fb, batch=batch, q0=q, v0=v)
    oracles = [pr.make_oracle(fb, q0=q[b], v0=v[b]) for b in range(batch)]
    n0 = len(solver.chain())
    seen_new = False
    for t in (0.0, 0.2, 0.41):
        for o in oracles:
            pr.set_references(o, t)
        rcs = [o.update_solution(t, q[b], v[b]) for b, o in enumerate(oracles)]
        assert not any(rcs), (t, rcs)
        solver.updateSolution(t, q, v)
        ch = solver.chain()
        assert chain_signature(ch) == chain_signature(oracles[0].chain()), t
        if len(ch) > n0:
            seen_new = True
            e = [k for k, c in enumerate(ch) if c["kind"] == fb.K_AUX][-1]
            slack = np.asarray(solver.get(e, "slack"))
            assert np.all(slack[:, 48:72] > 0), "torque-limit constraints of the entering aux stage (acceleration level) are disabled"
            assert np.array_equal(slack, fb.batch_get(oracles, e, "slack"))
            assert np.array_equal(np.asarray(solver.get(e - 1, "slack")), fb.batch_get(oracles, e - 1, "slack"))
        assert compare_batch(oracles, solver, fb, SOL + DIR) == [], t
    assert seen_new


def run_mpc_ticks(lib, fb, batch, ticks=(0.0, 0.2, 0.45, 0.52, 0.6, 0.8), iterations=2):
    """idocp_b200.BatchedMPC: >= 5 control ticks of a batch of MPC loops -- two Newton iterations per tick
    from the previous tick's solution, the plant moves to the second stage of the solution, the first phase is popped
    when its switching time has passed -- against the oracle driven with the same explicit calls, every instance bit for
    bit (first control input, feedback gain, the whole iterate)."""
    import idocp_b200 as I
    pr = ap.TrottingProblem()
    q, v = anymal_states(pr, batch, 20240004)
    solver = ap.make_product_solver(pr, lib, fb, batch=batch, q0=q, v0=v)
    oracles = [pr.make_oracle(fb, q0=q[b], v0=v[b]) for b in range(batch)]
    mpc = I.BatchedMPC(solver, iterations=iterations)
    pops = 0
    for t in ticks:
        u0, (Kq, Kv) = mpc.tick(t, q, v, with_gain=True)
        cs = oracles[0].cs
        while cs.counts()[1] + cs.counts()[2] > 0:
            first = min(([cs.impulse(0)[2]] if cs.counts()[1] else []) + ([cs.lift_time(0)] if cs.counts()[2] else []))
            if first > t:
                break
            for o in oracles:
                o.cs.pop_front()
            pops += 1
        for _ in range(iterations):
            for o in oracles:
                pr.set_references(o, t)
            rcs = [o.update_solution(t, q[b], v[b]) for b, o in enumerate(oracles)]
            assert all(rc == 0 for rc in rcs), t
        assert chain_signature(solver.chain()) == chain_signature(oracles[0].chain()), t
        assert np.array_equal(u0, fb.batch_get(oracles, 0, "u")), t
        K = fb.batch_get(oracles, 0, "K").reshape(batch, 12, 36)
        assert np.array_equal(Kq, K[:, :, :18]) and np.array_equal(Kv, K[:, :, 18:]), t
        assert compare_batch(oracles, solver, fb, SOL) == [], t
        q, v = fb.batch_get(oracles, 1, "q"), fb.batch_get(oracles, 1, "v")
    assert mpc.popped == pops and pops >= 1

"""Host-side contact schedule: the product (include/idocp_b200/hybrid.hpp through
the C-ABI) against the oracle restatement (oracle/hybrid_oracle.c) and against the scenarios of the reference's
own tests (test/hybrid/contact_sequence_test.cpp, discrete_event_test.cpp, ocp_discretizer_test.cpp,
test/robot/impulse_status_test.cpp).  No GPU involved: the schedule stays on the host."""
import numpy as np
import pytest

import idocp_b200 as I

MIN_DT = np.sqrt(np.finfo(float).eps)


@pytest.fixture(scope="module")
def H(oracle):
    import hybrid_py
    return hybrid_py


def _random_status(rng, n, different_from=None):
    while True:
        s = rng.integers(0, 2, n)
        if different_from is None or not np.array_equal(s, different_from):
            return s


def _pair(emu_lib, H, n=4, max_events=5):
    return I.ContactSequence(n, max_events, lib=emu_lib), H.ContactSequence(n, max_events)


def _build(rng, emu_lib, H, t, dt, n=4, max_events=5, on_grid=False, with_points=False):
    """createContactSequence / createContactSequenceOnGrid of ocp_discretizer_test.cpp:44-86 on both sides."""
    cs, ocs = _pair(emu_lib, H, n, max_events)
    pre = _random_status(rng, n)
    pts = rng.standard_normal((n, 3)) if with_points else None
    cs.setContactStatusUniformly(pre, pts)
    ocs.set_uniform(pre, pts)
    period = 3 * dt
    for i in range(max_events):
        post = _random_status(rng, n, different_from=pre)
        if on_grid:
            te = t + (i + 1) * period + MIN_DT * rng.uniform(-1, 1)
        else:
            te = t + i * period + dt * rng.uniform(0.05, 0.95)
        pts = rng.standard_normal((n, 3)) if with_points else None
        cs.push_back(post, te, pts)
        assert ocs.push_back(post, te, pts) == 0
        pre = post
    return cs, ocs


def _same_sequence(cs, ocs):
    assert (cs.numContactPhases(), cs.numImpulseEvents(), cs.numLiftEvents()) == ocs.counts()
    for k in range(cs.numContactPhases()):
        a, p = cs.contactStatus(k)
        oa, op = ocs.phase(k)
        assert np.array_equal(a, oa) and np.array_equal(p, op)
    for k in range(cs.numImpulseEvents()):
        a, p = cs.impulseStatus(k)
        oa, op, ot = ocs.impulse(k)
        assert np.array_equal(a, oa) and np.array_equal(p, op) and cs.impulseTime(k) == ot
    for k in range(cs.numLiftEvents()):
        assert cs.liftTime(k) == ocs.lift_time(k)


def _same_discretization(d, od):
    assert bool(od.well_defined) and (d.N(), d.N_impulse(), d.N_lift()) == (od.N, od.N_impulse, od.N_lift)
    for i in range(d.N() + 1):
        assert d.t(i) == od.t[i] and d.contactPhase(i) == od.contact_phase[i]
        if i < d.N():
            assert d.dt(i) == od.dt[i]
            assert d.impulseIndexAfterTimeStage(i) == od.impulse_after[i]
            assert d.liftIndexAfterTimeStage(i) == od.lift_after[i]
    for k in range(d.N_impulse()):
        assert (d.timeStageBeforeImpulse(k), d.t_impulse(k), d.dt_aux(k)) == (od.stage_before_impulse[k], od.t_impulse[k], od.dt_aux[k])
    for k in range(d.N_lift()):
        assert (d.timeStageBeforeLift(k), d.t_lift(k), d.dt_lift(k)) == (od.stage_before_lift[k], od.t_lift[k], od.dt_lift[k])


def test_discrete_event_classification(emu_lib, H):
    """discrete_event_test.cpp: impulse iff a contact becomes active (possibly lifting others), lift otherwise; the
    impulse status marks exactly the contacts that switch on (impulse_status_test.cpp setActivity)."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        cs, ocs = _pair(emu_lib, H)
        pre = _random_status(rng, 4)
        post = _random_status(rng, 4, different_from=pre)
        cs.setContactStatusUniformly(pre)
        ocs.set_uniform(pre)
        cs.push_back(post, 0.3)
        assert ocs.push_back(post, 0.3) == 0
        is_impulse = bool(np.any((pre == 0) & (post == 1)))
        assert cs.numImpulseEvents() == int(is_impulse) and cs.numLiftEvents() == int(not is_impulse)
        if is_impulse:
            assert np.array_equal(cs.impulseStatus(0)[0], ((pre == 0) & (post == 1)).astype(np.int32))
        _same_sequence(cs, ocs)


def test_contact_sequence_push_pop_update(emu_lib, H):
    """contact_sequence_test.cpp: push_back / pop_back / pop_front / updateImpulseTime / updateLiftTime /
    setContactPoints keep product and oracle in lock-step; the reference's error conditions are reported."""
    rng = np.random.default_rng(1)
    for trial in range(20):
        cs, ocs = _build(rng, emu_lib, H, t=0.1, dt=0.05, with_points=True)
        _same_sequence(cs, ocs)
        # move the first event a little, inside its admissible interval
        if cs.numImpulseEvents() and cs.numLiftEvents():
            first_is_impulse = cs.impulseTime(0) < cs.liftTime(0)
        else:
            first_is_impulse = cs.numImpulseEvents() > 0
        t0 = cs.impulseTime(0) if first_is_impulse else cs.liftTime(0)
        if first_is_impulse:
            cs.updateImpulseTime(0, t0 + 0.001)
        else:
            cs.updateLiftTime(0, t0 + 0.001)
        assert ocs.update_event_time(first_is_impulse, 0, t0 + 0.001) == 0
        _same_sequence(cs, ocs)
        pts = rng.standard_normal((4, 3))
        phase = int(rng.integers(0, cs.numContactPhases()))
        cs.setContactPoints(phase, pts)
        assert ocs.set_contact_points(phase, pts) == 0
        _same_sequence(cs, ocs)
        for op in rng.integers(0, 2, 7):
            (cs.pop_back, cs.pop_front)[op]()
            (ocs.pop_back, ocs.pop_front)[op]()
            _same_sequence(cs, ocs)
        assert cs.numContactPhases() == 1 and not cs.contactStatus(0)[0].any()   # back to the default status
    # error paths (contact_sequence.hxx:58-89, :157-187)
    cs, ocs = _pair(emu_lib, H, max_events=2)
    cs.setContactStatusUniformly([1, 1, 0, 0])
    with pytest.raises(I.Idocp_b200Error, match="existDiscreteEvent"):
        cs.push_back([1, 1, 0, 0], 0.1)
    cs.push_back([1, 1, 1, 0], 0.2)
    with pytest.raises(I.Idocp_b200Error, match="must be larger than the last event time"):
        cs.push_back([1, 1, 1, 1], 0.2)
    cs.push_back([1, 1, 1, 1], 0.3)
    with pytest.raises(I.Idocp_b200Error, match="exceeds predefined max_num_events=2"):
        cs.push_back([0, 1, 1, 1], 0.4)
    with pytest.raises(I.Idocp_b200Error, match="numLiftEvents\\(\\) must be positive"):
        cs.updateLiftTime(0, 0.25)
    with pytest.raises(I.Idocp_b200Error, match="must be less than numImpulseEvents\\(\\)=2"):
        cs.updateImpulseTime(2, 0.25)
    with pytest.raises(I.Idocp_b200Error, match="must be larger than event_time_"):
        cs.updateImpulseTime(1, 0.15)
    with pytest.raises(I.Idocp_b200Error, match="max_num_events must be positive"):
        I.ContactSequence(4, 0, lib=emu_lib)


def test_discretizer_constructor_state(emu_lib, H):
    """ocp_discretizer_test.cpp testConstructor: no events -> the plain grid."""
    cs, ocs = _pair(emu_lib, H)
    d = I.OCPDiscretizer(1.0, 20, lib=emu_lib)
    assert d.discretizeOCP(cs, 0.0)
    assert (d.N(), d.N_impulse(), d.N_lift(), d.N_all()) == (20, 0, 0, 21)
    assert all(d.contactPhase(i) == 0 for i in range(21))
    assert not any(d.isTimeStageBeforeImpulse(i) or d.isTimeStageBeforeLift(i) for i in range(20))
    kinds = [s.kind for s in d.stages()]
    assert kinds == [0] * 20 + [4]


@pytest.mark.parametrize("seed", range(12))
def test_discretize_ocp_generic_events(emu_lib, H, seed):
    """ocp_discretizer_test.cpp testDiscretizeOCP: one event every three grid intervals at a random offset."""
    rng = np.random.default_rng(100 + seed)
    N, T, max_events = 20, 1.0, 5
    dt = T / N
    t = abs(rng.uniform(-1, 1))
    cs, ocs = _build(rng, emu_lib, H, t, dt)
    d = I.OCPDiscretizer(T, N, lib=emu_lib)
    assert d.discretizeOCP(cs, t)
    _same_discretization(d, ocs.discretize(T, N, t))
    # the reference test's own expectations
    assert d.N() == N and d.N_impulse() == cs.numImpulseEvents() and d.N_lift() == cs.numLiftEvents()
    before = []
    for k in range(d.N_impulse()):
        ti = cs.impulseTime(k)
        st = int(np.floor((ti - t) / dt))
        before.append(st)
        assert d.timeStageBeforeImpulse(k) == st and d.t_impulse(k) == ti
        assert d.dt(st) == pytest.approx(ti - st * dt - t, rel=1e-12) and d.dt(st) + d.dt_aux(k) == pytest.approx(dt, rel=1e-12)
        assert d.impulseIndexAfterTimeStage(st) == k
    for k in range(d.N_lift()):
        tl = cs.liftTime(k)
        st = int(np.floor((tl - t) / dt))
        before.append(st)
        assert d.timeStageBeforeLift(k) == st and d.t_lift(k) == tl
        assert d.dt(st) + d.dt_lift(k) == pytest.approx(dt, rel=1e-12)
        assert d.liftIndexAfterTimeStage(st) == k
    before = sorted(before) + [N + 1]
    phase = 0
    for i in range(N + 1):
        assert d.contactPhase(i) == phase
        assert d.t(i) == pytest.approx(t + i * dt, rel=1e-12)
        if i == before[phase]:
            phase += 1
    # flattened schedule: Riccati visiting order, N_all rows
    st = d.stages()
    assert len(st) == d.N_all() and st[-1].kind == 4 and st[-1].t == t + T
    assert [s.index for s in st if s.kind == 0] == list(range(N))
    for pos, s in enumerate(st):
        if s.kind == 1:      # impulse: preceded by its grid stage, followed by its aux stage, no duration
            assert st[pos - 1].kind == 0 and st[pos - 1].index == d.timeStageBeforeImpulse(s.index)
            assert st[pos + 1].kind == 2 and st[pos + 1].index == s.index and s.dt == 0.0 and s.constraint_stage == -1
            assert st[pos + 1].dt == d.dt_aux(s.index) and st[pos + 1].constraint_stage == 0
            assert s.contact_phase == st[pos - 1].contact_phase + 1
            if pos >= 2 and st[pos - 2].kind == 0:   # switching constraint two stages ahead of the touch-down
                assert st[pos - 2].before_impulse == 1 and st[pos - 2].switching_impulse == s.index
        if s.kind == 3:
            assert st[pos - 1].kind == 0 and st[pos - 1].index == d.timeStageBeforeLift(s.index)
            assert s.dt == d.dt_lift(s.index) and s.constraint_stage == 0
    assert sum(s.before_impulse for s in st) <= d.N_impulse()
    assert abs(sum(s.dt for s in st) - T) < 1e-12            # the stage lengths tile the horizon


@pytest.mark.parametrize("seed", range(12))
def test_discretize_ocp_events_on_grid(emu_lib, H, seed):
    """ocp_discretizer_test.cpp testDiscretizeOCPOnGrid: events within sqrt(eps) of a grid point (either side)
    merge that grid stage away: N shrinks by the number of events."""
    rng = np.random.default_rng(200 + seed)
    N, T, max_events = 20, 1.0, 5
    dt = T / N
    t = abs(rng.uniform(-1, 1))
    cs, ocs = _build(rng, emu_lib, H, t, dt, on_grid=True)
    d = I.OCPDiscretizer(T, N, lib=emu_lib)
    assert d.discretizeOCP(cs, t)
    _same_discretization(d, ocs.discretize(T, N, t))
    assert d.N() == N - max_events
    ti = t
    for i in range(d.N()):
        assert abs(d.dt(i) - dt) <= MIN_DT and abs(d.t(i) - ti) <= MIN_DT
        ti += dt
        if d.isTimeStageBeforeImpulse(i) or d.isTimeStageBeforeLift(i):
            ti += dt
    assert d.t(d.N()) == t + T
    assert abs(sum(s.dt for s in d.stages()) - T) < 1e-6


def test_anymal_example_schedules(emu_lib, H):
    """Stage counts of SURVEY Appendix C: trotting (T = 1.55, N = 30: lift at 0.5, touch-down + lift at 1.0 and 1.5
    -> 36 stages) with the contact pattern of examples/anymal/anymal_trotting.cpp:144-177."""
    cs, ocs = _pair(emu_lib, H, n=4, max_events=3)
    standing, lfrh, rflh = [1, 1, 1, 1], [0, 1, 1, 0], [1, 0, 0, 1]
    cs.setContactStatusUniformly(standing)
    ocs.set_uniform(standing)
    for status, te in ((lfrh, 0.5), (rflh, 1.0), (lfrh, 1.5)):
        cs.push_back(status, te)
        assert ocs.push_back(status, te) == 0
    assert (cs.numImpulseEvents(), cs.numLiftEvents()) == (2, 1)
    d = I.OCPDiscretizer(1.55, 30, lib=emu_lib)
    assert d.discretizeOCP(cs, 0.0)
    _same_discretization(d, ocs.discretize(1.55, 30, 0.0))
    assert d.N_all() == 31 + 2 * 2 + 1 == 36 and len(d.stages()) == 36


def test_ill_defined_schedules_are_reported(emu_lib, H):
    """The reference only asserts isWellDefined() in Debug builds; here the flag is returned: two events in one grid
    interval, an event before t, impulses after consecutive stages.  Events beyond the horizon are left unscheduled."""
    d = I.OCPDiscretizer(1.0, 20, lib=emu_lib)
    cs, ocs = _pair(emu_lib, H)
    cs.setContactStatusUniformly([1, 1, 1, 1])
    cs.push_back([0, 1, 1, 1], 0.21)
    cs.push_back([0, 0, 1, 1], 0.23)
    assert not d.discretizeOCP(cs, 0.0)          # same interval [0.20, 0.25)
    cs, ocs = _pair(emu_lib, H)
    cs.setContactStatusUniformly([1, 1, 1, 1])
    cs.push_back([0, 1, 1, 1], 0.21)
    assert not d.discretizeOCP(cs, 0.3)          # event in the past
    cs, ocs = _pair(emu_lib, H)
    cs.setContactStatusUniformly([0, 0, 1, 1])
    cs.push_back([1, 0, 1, 1], 0.21)
    cs.push_back([1, 1, 1, 1], 0.27)
    assert not d.discretizeOCP(cs, 0.0)          # impulses after stages 4 and 5
    cs, ocs = _pair(emu_lib, H)
    cs.setContactStatusUniformly([1, 1, 1, 1])
    cs.push_back([0, 1, 1, 1], 0.21)
    cs.push_back([1, 1, 1, 1], 1.7)              # beyond t + T: not part of this discretisation
    assert d.discretizeOCP(cs, 0.0) and (d.N_impulse(), d.N_lift()) == (0, 1)

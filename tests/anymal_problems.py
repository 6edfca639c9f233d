"""The reference's ANYmal example problems as plain data (used by the oracle tests and the GPU parity tests).

anymal_trotting: examples/anymal/anymal_trotting.cpp:30-196 (cost weights, constraints, contact schedule, initial
guess); the time-varying configuration reference is TrottingConfigurationSpaceCost::update_q_ref
(include/idocp/cost/trotting_configuration_space_cost.hpp:126-164), sampled at the time of every stage."""
import math

import numpy as np

import fb_py
import hybrid_py

Q_STANDING = np.array([0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0])
TOTAL_WEIGHT = 30.475397462000004 * 9.81


def trotting_q_ref(t, t_start, t_period, q_standing, step_length, swing):
    """swing: dict front_swing_knee, hip_swing_knee, front_stance_knee, hip_stance_knee (others unused upstream)."""
    q = np.array(q_standing, dtype=float)
    if t > t_start:
        tau = t - t_start
        steps = math.floor(tau / t_period)
        tau_step = tau - steps * t_period
        rate = tau_step / t_period
        sin2 = math.sin(0.5 * math.pi * rate)
        q[0] += (steps + rate) * step_length
        if steps % 2 == 0:
            q[9] -= sin2 * swing.get("front_swing_knee", 0.0)
            q[12] -= sin2 * swing.get("hip_stance_knee", 0.0)
            q[15] += sin2 * swing.get("front_stance_knee", 0.0)
            q[18] += sin2 * swing.get("hip_swing_knee", 0.0)
        else:
            q[9] += sin2 * swing.get("front_stance_knee", 0.0)
            q[12] += sin2 * swing.get("hip_swing_knee", 0.0)
            q[15] -= sin2 * swing.get("front_swing_knee", 0.0)
            q[18] -= sin2 * swing.get("hip_stance_knee", 0.0)
    return q


def standing_contact_points(fb):
    z = np.zeros(18)
    return np.stack([fb.contact(Q_STANDING, z, z, i, 0.05, np.zeros(3))["P"] for i in range(4)])


class TrottingProblem:
    """anymal_trotting.cpp with `steps` impulse phases (the shipped example: 2 steps, T = 1.55, N = 30)."""

    def __init__(self, steps=2):
        self.step_length, self.t_start, self.t_period = 0.15, 0.5, 0.5
        self.swing = dict(front_swing_knee=1.7, hip_swing_knee=1.7)
        self.T = self.t_start + steps * self.t_period + 0.05
        self.N = 10 + 10 * steps          # the example's table: (T, N) = (1.55, 30), (2.55, 50), (3.55, 70), ...
        self.steps = steps
        self.max_num_impulse = steps + 1
        p = fb_py.FbProblem()
        p.T, p.N, p.max_num_impulse = self.T, self.N, self.max_num_impulse
        qw = np.full(18, 10.0)
        vw = np.array([1.0] * 6 + [0.1] * 12)
        aw = np.array([0.1] * 6 + [0.01] * 12)
        for nm in ("q_weight", "qf_weight", "qi_weight"):
            p.set(nm, qw)
        for nm in ("v_weight", "vf_weight", "vi_weight"):
            p.set(nm, vw)
        p.set("a_weight", aw)
        p.set("dvi_weight", aw)
        p.set("f_weight", np.full(12, 0.001))
        p.set("fi_weight", np.full(12, 0.001))
        fref = np.tile([0, 0, TOTAL_WEIGHT / 4], 4)
        p.set("f_ref", fref)
        p.set("fi_ref", np.zeros(12))          # ContactForceCost::fi_ref_ stays zero (set_f_ref only touches f_ref_)
        p.set("q_min", np.full(12, -9.42))
        p.set("q_max", np.full(12, 9.42))
        p.set("v_max", np.full(12, 15.0))
        p.set("u_max", np.full(12, 80.0))
        p.mu, p.barrier, p.fraction_rate = 0.7, 1.0e-4, 0.995
        for c in range(8):
            p.enable[c] = 1
        self.problem = p
        self.v_ref = np.zeros(18)
        self.v_ref[0] = self.step_length / self.t_period
        self.q0 = Q_STANDING.copy()
        self.v0 = np.zeros(18)
        self.f_init = np.array([0, 0, 0.25 * TOTAL_WEIGHT])
        self.standing_points = None     # override of the contact points of the standing pose (default: the oracle's FK)

    def contact_sequence(self, fb):
        cs = hybrid_py.ContactSequence(4, self.max_num_impulse + 2)
        pts = standing_contact_points(fb) if self.standing_points is None else np.array(self.standing_points, dtype=float)
        cs.set_uniform([1, 1, 1, 1], pts)
        cs.push_back([0, 1, 1, 0], self.t_start, pts)
        pts = pts.copy()
        pts[0, 0] += 0.5 * self.step_length
        pts[3, 0] += 0.5 * self.step_length
        cs.push_back([1, 0, 0, 1], self.t_start + self.t_period, pts)
        for i in range(2, self.steps + 1):
            pts = pts.copy()
            if i % 2 == 0:
                pts[1, 0] += self.step_length
                pts[2, 0] += self.step_length
                cs.push_back([0, 1, 1, 0], self.t_start + i * self.t_period, pts)
            else:
                pts[0, 0] += self.step_length
                pts[3, 0] += self.step_length
                cs.push_back([1, 0, 0, 1], self.t_start + i * self.t_period, pts)
        return cs

    def q_ref(self, t):
        return trotting_q_ref(t, self.t_start, self.t_period, Q_STANDING, self.step_length, self.swing)

    def make_oracle(self, fb, t=0.0, q0=None, v0=None):
        """OCPSolver construction + initial guess + initConstraints exactly as the example's main()."""
        cs = self.contact_sequence(fb)
        ocp = fb.FbOCP(self.problem, cs)
        ocp.set_solution("q", self.q0 if q0 is None else q0)
        ocp.set_solution("v", self.v0 if v0 is None else v0)
        ocp.set_solution("f", self.f_init)
        self.set_references(ocp, t)
        ocp.init_constraints(t)
        return ocp

    def set_references(self, ocp, t):
        ocp.discretize(t)
        for el in ocp.chain():
            kind = fb_py.K_GRID if el["kind"] == fb_py.K_TERMINAL else el["kind"]
            ocp.set_reference(kind, el["index"], self.q_ref(el["t"]), self.v_ref)


def with_nonlinear_cones_and_acceleration_limits(pr, cones=True, a_limit=9.0):
    """SURVEY 8(f3) components on an ANYmal problem: FrictionCone + ImpulseFrictionCone (src/constraints/friction_cone.cpp,
    impulse_friction_cone.cpp) instead of the linearised cones, JointAcceleration{Lower,Upper}Limit
    (joint_acceleration_*_limit.cpp) with amin = -a_limit, amax = +a_limit (the unconstrained trot peaks at 18 rad/s^2)."""
    p = pr.problem
    p.cone_nonlinear[0] = p.cone_nonlinear[1] = 1 if cones else 0
    if a_limit is not None:
        p.enable_acc[0] = p.enable_acc[1] = 1
        p.set("a_min", np.full(12, -a_limit))
        p.set("a_max", np.full(12, a_limit))
    return pr


class JumpingProblem(TrottingProblem):
    """A flight phase: all feet leave at t_lift (a lift stage, dimf = 0 afterwards) and touch down at t_land
    (an impulse with four contacts), in the style of examples/anymal/anymal_jumping.cpp."""

    def __init__(self, jump_length=0.2, t_lift=0.42, t_land=0.73, T=1.1, N=22):
        super().__init__(steps=2)
        self.T, self.N, self.max_num_impulse = T, N, 2
        self.problem.T, self.problem.N, self.problem.max_num_impulse = T, N, 2
        self.jump_length, self.t_lift, self.t_land = jump_length, t_lift, t_land
        self.v_ref = np.zeros(18)

    def contact_sequence(self, fb):
        cs = hybrid_py.ContactSequence(4, 4)
        pts = standing_contact_points(fb)
        cs.set_uniform([1, 1, 1, 1], pts)
        cs.push_back([0, 0, 0, 0], self.t_lift, pts)
        pts = pts.copy()
        pts[:, 0] += self.jump_length
        cs.push_back([1, 1, 1, 1], self.t_land, pts)
        return cs

    def q_ref(self, t):
        q = Q_STANDING.copy()
        if t > self.t_land:
            q[0] += self.jump_length
        elif t > self.t_lift:
            q[0] += self.jump_length * (t - self.t_lift) / (self.t_land - self.t_lift)
        return q


class StandingBenchmarkProblem(TrottingProblem):
    """examples/anymal/ocp_benchmark.cpp:26-125: four-foot stance, ConfigurationSpaceCost (constant reference) +
    ContactForceCost (f_ref = (0, 0, 70)), six joint limits from the URDF and the nonlinear FrictionCone(mu = 0.7);
    T = 0.5, N = 20, max_num_impulse = 4, 10 iterations."""

    def __init__(self):
        super().__init__(steps=2)
        self.T, self.N, self.max_num_impulse = 0.5, 20, 4
        p = self.problem
        p.T, p.N, p.max_num_impulse = 0.5, 20, 4
        for nm in ("q_weight", "qf_weight"):
            p.set(nm, np.full(18, 10.0))
        for nm in ("v_weight", "vf_weight"):
            p.set(nm, np.full(18, 1.0))
        p.set("a_weight", np.full(18, 0.01))
        for nm in ("qi_weight", "vi_weight", "dvi_weight", "fi_weight", "fi_ref"):    # never set by the example: zero
            p.set(nm, np.zeros(len(getattr(p, nm))))
        p.set("f_weight", np.full(12, 0.001))
        p.set("f_ref", np.tile([0, 0, 70.0], 4))
        # the joint limits of the URDF (the trotting example overrides nothing either; the product's Robot holds the same
        # numbers: idocp_b200_fb_problem_default)
        self.urdf_limits = True
        p.mu = 0.7
        for c in range(8):
            p.enable[c] = 1 if c < 7 else 0        # no impulse cone in the example
        p.cone_nonlinear[0] = 1
        self.v_ref = np.zeros(18)

    def contact_sequence(self, fb):
        cs = hybrid_py.ContactSequence(4, 2 * self.max_num_impulse + 2)
        pts = standing_contact_points(fb) if self.standing_points is None else np.array(self.standing_points, dtype=float)
        cs.set_uniform([1, 1, 1, 1], pts)
        return cs

    def q_ref(self, t):
        return Q_STANDING.copy()


def make_product_solver(pr, lib, fb, batch, q0=None, v0=None, t=0.0, devices=None):
    """The product's OCPSolver (idocp_b200/ocp_solver.py over the C-ABI) set up as the example's main() does; q0 / v0
    may be (batch, dim) arrays of per-instance initial states (also used as the initial guess, like the example)."""
    import ctypes as C
    import idocp_b200 as I
    p = I.FbProblem()
    assert C.sizeof(p) == C.sizeof(pr.problem)
    C.memmove(C.byref(p), C.byref(pr.problem), C.sizeof(p))
    solver = I.OCPSolver(p, batch, q_ref=lambda tt: (pr.q_ref(tt), pr.v_ref), lib=lib, max_num_events=60, devices=devices)
    ocs = pr.contact_sequence(fb)
    n_phases = ocs.counts()[0]
    a, pts = ocs.phase(0)
    solver.setContactStatusUniformly(a, pts)
    # replay the oracle-side schedule through the product's own ContactSequence
    events = []
    ni = nl = 0
    for k in range(1, n_phases):
        a, pts = ocs.phase(k)
        events.append((a, pts))
    times = sorted([ocs.impulse(i)[2] for i in range(ocs.counts()[1])] + [ocs.lift_time(i) for i in range(ocs.counts()[2])])
    for (a, pts), tt in zip(events, times):
        solver.pushBackContactStatus(a, tt, pts)
    solver.setSolution("q", pr.q0 if q0 is None else q0)
    solver.setSolution("v", pr.v0 if v0 is None else v0)
    solver.setSolution("f", pr.f_init)
    solver.initConstraints(t)
    return solver


class RunningProblem(TrottingProblem):
    """examples/anymal/anymal_running.cpp:28-229: T = 7, N = 240, a running gait with flight phases (26 impulses,
    14 lifts), TimeVaryingConfigurationSpaceCost (cost/time_varying_configuration_space_cost.hpp:98-118: the reference
    moves with v_ref between t_begin and t_end) and a contact-force cost; `steps` shortens the gait for the tests
    (the example uses 10; T and N shrink with it so that the schedule still ends in a four-foot stance)."""

    def __init__(self, steps=10, fb=None):
        import fb_py as _fb
        self.fb = fb or _fb
        self.stride, self.additive_stride_hip, self.t_start = 0.4, 0.2, 1.0
        self.t_front_swing, self.t_front_hip_swing, self.t_hip_swing = 0.135, 0.05, 0.165
        self.t_period = self.t_front_swing + self.t_front_hip_swing + self.t_hip_swing
        self.steps = steps
        self.max_num_impulse = (steps + 3) * 2
        if steps == 10:
            self.T, self.N = 7.0, 240
        else:   # same grid spacing, horizon ending 0.6 s after the last touch-down
            t_last = self.t_start + 0.30 + 0.34 + steps * self.t_period + 0.35
            self.N = int(math.ceil((t_last + 0.6) / (7.0 / 240)))
            self.T = self.N * (7.0 / 240)
        p = fb_py.FbProblem()
        p.T, p.N, p.max_num_impulse = self.T, self.N, self.max_num_impulse
        qw = np.array([1.0] * 3 + [10.0] * 15)
        vw = np.array([0.01] * 3 + [0.1] * 15)
        aw = np.full(18, 0.01)
        for nm in ("q_weight", "qf_weight", "qi_weight"):
            p.set(nm, qw)
        for nm in ("v_weight", "vf_weight", "vi_weight"):
            p.set(nm, vw)
        p.set("a_weight", aw)
        p.set("dvi_weight", aw)
        p.set("f_weight", np.tile([0.1, 0.1, 1.0e-07], 4))
        p.set("fi_weight", np.tile([0.1, 0.1, 1.0e-07], 4))
        p.set("f_ref", np.tile([0, 0, 70.0], 4))
        p.set("fi_ref", np.zeros(12))
        p.set("q_min", np.full(12, -9.42))
        p.set("q_max", np.full(12, 9.42))
        p.set("v_max", np.full(12, 15.0))
        p.set("u_max", np.full(12, 80.0))
        p.mu, p.barrier, p.fraction_rate = 0.8, 1.0e-4, 0.995
        for c in range(8):
            p.enable[c] = 1
        self.problem = p
        self.q_begin = Q_STANDING.copy()
        self.q_begin[0] = -3.0
        self.v_ref_moving = np.zeros(18)
        self.v_ref_moving[0] = self.stride / self.t_period
        self.t_begin, self.t_end = self.t_start, self.t_start + (0.5 + steps) * self.t_period
        self.q_end = self.fb.integrate(self.q_begin, self.v_ref_moving, self.t_end - self.t_begin)
        self.q0 = self.q_begin.copy()
        self.v0 = np.zeros(18)
        self.f_init = np.array([0, 0, 0.25 * TOTAL_WEIGHT])
        self.standing_points = None
        self._v_ref_t = None

    def q_ref(self, t):
        if self.t_begin < t < self.t_end:
            self.v_ref = self.v_ref_moving
            return self.fb.integrate(self.q_begin, self.v_ref_moving, t - self.t_begin)
        self.v_ref = np.zeros(18)
        return self.q_begin.copy() if t <= self.t_begin else self.q_end.copy()

    def set_references(self, ocp, t):
        ocp.discretize(t)
        for el in ocp.chain():
            kind = fb_py.K_GRID if el["kind"] == fb_py.K_TERMINAL else el["kind"]
            q_ref = self.q_ref(el["t"])
            ocp.set_reference(kind, el["index"], q_ref, self.v_ref)

    def contact_sequence(self, fb):
        z = np.zeros(18)
        pts = np.stack([fb.contact(self.q_begin, z, z, i, 0.05, np.zeros(3))["P"] for i in range(4)])
        if self.standing_points is not None:
            pts = np.array(self.standing_points, dtype=float)
        cs = hybrid_py.ContactSequence(4, 64)
        ALL, FRONT_SWING, FLY, HIP_SWING = [1, 1, 1, 1], [0, 1, 0, 1], [0, 0, 0, 0], [1, 0, 1, 0]
        st, ah = self.stride, self.additive_stride_hip
        t0 = self.t_start
        cs.set_uniform(ALL, pts)
        cs.push_back(FRONT_SWING, t0, pts)
        cs.push_back(FLY, t0 + 0.125, pts)
        pts = pts + np.array([[0.25 * st], [0.25 * st + 0.5 * ah], [0.25 * st], [0.25 * st + 0.5 * ah]]) * np.array([1.0, 0, 0])
        cs.push_back(HIP_SWING, t0 + 0.125 + 0.05, pts)
        t_initial, t_initial2 = 0.125 + 0.05 + 0.125, 0.135 + 0.055 + 0.15
        cs.push_back(FRONT_SWING, t0 + t_initial, pts)
        cs.push_back(FLY, t0 + t_initial + 0.135, pts)
        pts = pts + np.array([[0.5 * st], [0.5 * st + 0.5 * ah], [0.5 * st], [0.5 * st + 0.5 * ah]]) * np.array([1.0, 0, 0])
        cs.push_back(HIP_SWING, t0 + t_initial + 0.135 + 0.055, pts)
        t_end_init = t0 + t_initial + t_initial2
        for i in range(self.steps):
            cs.push_back(FRONT_SWING, t_end_init + i * self.t_period, pts)
            cs.push_back(FLY, t_end_init + i * self.t_period + self.t_front_swing, pts)
            pts = pts + np.array([st, 0, 0])
            cs.push_back(HIP_SWING, t_end_init + i * self.t_period + self.t_front_swing + self.t_front_hip_swing, pts)
        tl = t_end_init + self.steps * self.t_period
        cs.push_back(FRONT_SWING, tl, pts)
        cs.push_back(FLY, tl + 0.15, pts)
        pts = pts + np.array([[0.5 * st], [0.5 * st - ah], [0.5 * st], [0.5 * st - ah]]) * np.array([1.0, 0, 0])
        cs.push_back(HIP_SWING, tl + 0.15 + 0.05, pts)
        cs.push_back(ALL, tl + 0.35, pts)
        return cs

"""The reference's ANYmal example problems as plain data for the product's OCPSolver (no oracle involved).

anymal_trotting: examples/anymal/anymal_trotting.cpp:30-196; anymal_running: examples/anymal/anymal_running.cpp:28-229.
Contact points come from the product's host-side Robot::updateFrameKinematics (idocp_b200_fb_contact_frame_positions)."""
import math

import numpy as np

from . import capi
from .ocp_solver import OCPSolver

Q_STANDING = np.array([0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0])
SEED_ANYMAL = 20240004


def _lib(lib):
    return lib or capi.default_library()


def contact_points(q, lib=None):
    lib = _lib(lib)
    out = np.zeros((4, 3))
    lib.check(lib.L.idocp_b200_fb_contact_frame_positions(capi.dptr(np.ascontiguousarray(q, dtype=np.float64)), capi.dptr(out)))
    return out


def _base_problem(lib, T, N, max_num_impulse, qw, vw, aw, fw, fref, mu):
    lib = _lib(lib)
    p = capi.FbProblem()
    lib.check(lib.L.idocp_b200_fb_problem_default(p))
    p.T, p.N, p.max_num_impulse = T, N, max_num_impulse

    def put(name, values):
        arr = getattr(p, name)
        for i, x in enumerate(np.asarray(values, dtype=float).ravel()):
            arr[i] = x
    for nm in ("q_weight", "qf_weight", "qi_weight"):
        put(nm, qw)
    for nm in ("v_weight", "vf_weight", "vi_weight"):
        put(nm, vw)
    put("a_weight", aw)
    put("dvi_weight", aw)
    put("f_weight", np.tile(fw, 4))
    put("fi_weight", np.tile(fw, 4))
    put("f_ref", np.tile(fref, 4))
    p.mu = mu
    for c in range(capi.FB_NUM_CONSTRAINTS):
        p.enable[c] = 1
    return p


def splitmix_uniform(seed, index):
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def anymal_initial_states(first_instance, count, q_nominal=Q_STANDING, seed=SEED_ANYMAL):
    """base position +-0.01 m, base orientation exp(+-0.02 rad axis-angle), joints +-0.02 rad around
    the nominal pose, v0 = 0.1 U(-1,1); counter-based splitmix64, element index = instance * 36 + j."""
    inst = np.arange(first_instance, first_instance + count, dtype=np.uint64)[:, None]
    j = np.arange(36, dtype=np.uint64)[None, :]
    u = 2.0 * splitmix_uniform(seed, inst * np.uint64(36) + j) - 1.0
    q = np.tile(np.asarray(q_nominal, dtype=float), (count, 1))
    q[:, :3] += 0.01 * u[:, :3]
    w = 0.02 * u[:, 3:6]
    th = np.linalg.norm(w, axis=1, keepdims=True)
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.maximum(th, 1e-300), 0.5)
    dq = np.concatenate([k * w, np.cos(half)], axis=1)            # (x, y, z, w) of exp(w)
    x0, y0, z0, w0 = (q[:, 3 + i] for i in range(4))
    x1, y1, z1, w1 = (dq[:, i] for i in range(4))
    q[:, 3] = w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1
    q[:, 4] = w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1
    q[:, 5] = w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1
    q[:, 6] = w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1
    q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    q[:, 7:] += 0.02 * u[:, 6:18]
    v = 0.1 * u[:, 18:36]
    return np.ascontiguousarray(q), np.ascontiguousarray(v)


class AnymalTrotting:
    """examples/anymal/anymal_trotting.cpp (steps = 2: T = 1.55, N = 30, three events)."""
    name = "anymal_trotting"

    def __init__(self, steps=2, lib=None):
        self.lib = _lib(lib)
        self.step_length, self.t_start, self.t_period = 0.15, 0.5, 0.5
        self.front_swing_knee = self.hip_swing_knee = 1.7
        self.T = self.t_start + steps * self.t_period + 0.05
        self.N = 10 + 10 * steps
        self.steps = steps
        self.max_num_impulse = steps + 1
        weight = float(self.lib.L.idocp_b200_fb_total_weight())
        self.problem = _base_problem(self.lib, self.T, self.N, self.max_num_impulse, np.full(18, 10.0),
                                     np.array([1.0] * 6 + [0.1] * 12), np.array([0.1] * 6 + [0.01] * 12),
                                     [0.001, 0.001, 0.001], [0, 0, weight / 4], 0.7)
        self.v_ref = np.zeros(18)
        self.v_ref[0] = self.step_length / self.t_period
        self.q_nominal = Q_STANDING.copy()
        self.f_init = np.array([0, 0, 0.25 * weight])

    def q_ref(self, t):
        q = Q_STANDING.copy()
        if t > self.t_start:
            tau = t - self.t_start
            steps = math.floor(tau / self.t_period)
            rate = (tau - steps * self.t_period) / self.t_period
            sin2 = math.sin(0.5 * math.pi * rate)
            q[0] += (steps + rate) * self.step_length
            if steps % 2 == 0:
                q[9] -= sin2 * self.front_swing_knee
                q[18] += sin2 * self.hip_swing_knee
            else:
                q[12] += sin2 * self.hip_swing_knee
                q[15] -= sin2 * self.front_swing_knee
        return q, self.v_ref

    def schedule(self):
        pts = contact_points(Q_STANDING, self.lib)
        out = [([1, 1, 1, 1], None, pts)]
        out.append(([0, 1, 1, 0], self.t_start, pts))
        pts = pts.copy()
        pts[0, 0] += 0.5 * self.step_length
        pts[3, 0] += 0.5 * self.step_length
        out.append(([1, 0, 0, 1], self.t_start + self.t_period, pts))
        for i in range(2, self.steps + 1):
            pts = pts.copy()
            if i % 2 == 0:
                pts[1, 0] += self.step_length
                pts[2, 0] += self.step_length
                out.append(([0, 1, 1, 0], self.t_start + i * self.t_period, pts))
            else:
                pts[0, 0] += self.step_length
                pts[3, 0] += self.step_length
                out.append(([1, 0, 0, 1], self.t_start + i * self.t_period, pts))
        return out


class AnymalRunning:
    """examples/anymal/anymal_running.cpp (T = 7, N = 240, 26 impulses, 14 lifts, flight phases)."""
    name = "anymal_running"

    def __init__(self, lib=None):
        self.lib = _lib(lib)
        self.stride, self.additive_stride_hip, self.t_start = 0.4, 0.2, 1.0
        self.t_front_swing, self.t_front_hip_swing, self.t_hip_swing = 0.135, 0.05, 0.165
        self.t_period = self.t_front_swing + self.t_front_hip_swing + self.t_hip_swing
        self.steps = 10
        self.T, self.N, self.max_num_impulse = 7.0, 240, 26
        weight = float(self.lib.L.idocp_b200_fb_total_weight())
        self.problem = _base_problem(self.lib, self.T, self.N, self.max_num_impulse, np.array([1.0] * 3 + [10.0] * 15),
                                     np.array([0.01] * 3 + [0.1] * 15), np.full(18, 0.01), [0.1, 0.1, 1.0e-07], [0, 0, 70.0], 0.8)
        self.q_nominal = Q_STANDING.copy()
        self.q_nominal[0] = -3.0
        self.v_moving = np.zeros(18)
        self.v_moving[0] = self.stride / self.t_period
        self.t_begin, self.t_end = self.t_start, self.t_start + (0.5 + self.steps) * self.t_period
        self.f_init = np.array([0, 0, 0.25 * weight])

    def q_ref(self, t):
        # TimeVaryingConfigurationSpaceCost::set_q_ref / v_ref (time_varying_configuration_space_cost.hpp:98-118); the
        # base orientation of q_begin is the identity, so integrate(q_begin, v_ref, dt) is a translation along x
        q = self.q_nominal.copy()
        if self.t_begin < t < self.t_end:
            q[0] += self.v_moving[0] * (t - self.t_begin)
            return q, self.v_moving
        if t >= self.t_end:
            q[0] += self.v_moving[0] * (self.t_end - self.t_begin)
        return q, np.zeros(18)

    def schedule(self):
        pts = contact_points(self.q_nominal, self.lib)
        ALL, FRONT_SWING, FLY, HIP_SWING = [1, 1, 1, 1], [0, 1, 0, 1], [0, 0, 0, 0], [1, 0, 1, 0]
        st, ah, t0 = self.stride, self.additive_stride_hip, self.t_start

        def shifted(p, dx):
            p = p.copy()
            p[:, 0] += np.asarray(dx, dtype=float)
            return p
        out = [(ALL, None, pts), (FRONT_SWING, t0, pts), (FLY, t0 + 0.125, pts)]
        pts = shifted(pts, [0.25 * st, 0.25 * st + 0.5 * ah, 0.25 * st, 0.25 * st + 0.5 * ah])
        out.append((HIP_SWING, t0 + 0.125 + 0.05, pts))
        t_initial, t_initial2 = 0.125 + 0.05 + 0.125, 0.135 + 0.055 + 0.15
        out += [(FRONT_SWING, t0 + t_initial, pts), (FLY, t0 + t_initial + 0.135, pts)]
        pts = shifted(pts, [0.5 * st, 0.5 * st + 0.5 * ah, 0.5 * st, 0.5 * st + 0.5 * ah])
        out.append((HIP_SWING, t0 + t_initial + 0.135 + 0.055, pts))
        t_end_init = t0 + t_initial + t_initial2
        for i in range(self.steps):
            out += [(FRONT_SWING, t_end_init + i * self.t_period, pts), (FLY, t_end_init + i * self.t_period + self.t_front_swing, pts)]
            pts = shifted(pts, [st] * 4)
            out.append((HIP_SWING, t_end_init + i * self.t_period + self.t_front_swing + self.t_front_hip_swing, pts))
        tl = t_end_init + self.steps * self.t_period
        out += [(FRONT_SWING, tl, pts), (FLY, tl + 0.15, pts)]
        pts = shifted(pts, [0.5 * st, 0.5 * st - ah, 0.5 * st, 0.5 * st - ah])
        out += [(HIP_SWING, tl + 0.15 + 0.05, pts), (ALL, tl + 0.35, pts)]
        return out


def make_solver(problem, batch, q0, v0, device=0, t=0.0, lib=None):
    """OCPSolver set up the way the example's main() does: schedule, initial guess (= the initial state of every
    instance), f = a quarter of the weight on every foot, initConstraints."""
    solver = OCPSolver(problem.problem, batch, q_ref=problem.q_ref, device=device, lib=lib or problem.lib, max_num_events=60)
    sched = problem.schedule()
    solver.setContactStatusUniformly(sched[0][0], sched[0][2])
    for active, time, pts in sched[1:]:
        solver.pushBackContactStatus(active, time, pts)
    solver.setSolution("q", q0)
    solver.setSolution("v", v0)
    solver.setSolution("f", problem.f_init)
    solver.initConstraints(t)
    return solver

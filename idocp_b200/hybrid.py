"""Host-side contact schedule of the hybrid OCP through the C-ABI.

Mirrors idocp's ContactSequence (include/idocp/hybrid/contact_sequence.hpp) and OCPDiscretizer
(include/idocp/hybrid/ocp_discretizer.hpp); contact statuses are sequences of 0/1 flags, contact points
(max_point_contacts, 3) arrays.  All logic lives in include/idocp_b200/hybrid.hpp behind the C-ABI.
"""
import ctypes as C

import numpy as np

from . import capi

__all__ = ["ContactSequence", "OCPDiscretizer"]


def _flags(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
    if a.shape != (n,):
        raise ValueError("contact status must have %d entries" % n)
    return a


def _points(p, n):
    if p is None:
        return None
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
    if p.shape != (n, 3):
        raise ValueError("contact points must have shape (%d, 3)" % n)
    return p


class ContactSequence:
    def __init__(self, max_point_contacts, max_num_events, lib=None):
        self.lib = lib or capi.default_library()
        self.n = int(max_point_contacts)
        self.max_num_events = int(max_num_events)
        self._h = C.c_void_p()
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_create(self.n, self.max_num_events, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.L.idocp_b200_contact_sequence_destroy(self._h)
            self._h = None

    def _ip(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_int))

    def _dp(self, p):
        return p.ctypes.data_as(C.POINTER(C.c_double)) if p is not None else None

    def setContactStatusUniformly(self, is_active, contact_points=None):
        a, p = _flags(is_active, self.n), _points(contact_points, self.n)
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_set_uniform(self._h, self._ip(a), self._dp(p)))

    def push_back(self, is_active, event_time, contact_points=None):
        a, p = _flags(is_active, self.n), _points(contact_points, self.n)
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_push_back(self._h, self._ip(a), self._dp(p),
                                                                        float(event_time)))

    def pop_back(self):
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_pop_back(self._h))

    def pop_front(self):
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_pop_front(self._h))

    def updateImpulseTime(self, impulse_index, impulse_time):
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_update_event_time(self._h, 1, int(impulse_index),
                                                                                float(impulse_time)))

    def updateLiftTime(self, lift_index, lift_time):
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_update_event_time(self._h, 0, int(lift_index),
                                                                                float(lift_time)))

    def setContactPoints(self, contact_phase, contact_points):
        p = _points(contact_points, self.n)
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_set_contact_points(self._h, int(contact_phase),
                                                                                 self._dp(p)))

    def _counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def numContactPhases(self):
        return self._counts()[0]

    def numImpulseEvents(self):
        return self._counts()[1]

    def numLiftEvents(self):
        return self._counts()[2]

    def numDiscreteEvents(self):
        return self._counts()[0] - 1

    def contactStatus(self, contact_phase):
        a, p = np.zeros(self.n, dtype=np.int32), np.zeros((self.n, 3))
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_get_phase(self._h, int(contact_phase), self._ip(a),
                                                                        self._dp(p)))
        return a, p

    def impulseStatus(self, impulse_index):
        a, p, t = np.zeros(self.n, dtype=np.int32), np.zeros((self.n, 3)), C.c_double()
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_get_impulse(self._h, int(impulse_index), self._ip(a),
                                                                          self._dp(p), C.byref(t)))
        return a, p

    def impulseTime(self, impulse_index):
        t = C.c_double()
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_get_impulse(self._h, int(impulse_index), None, None,
                                                                          C.byref(t)))
        return t.value

    def liftTime(self, lift_index):
        t = C.c_double()
        self.lib.check(self.lib.L.idocp_b200_contact_sequence_get_lift_time(self._h, int(lift_index), C.byref(t)))
        return t.value


class OCPDiscretizer:
    """OCPDiscretizer(T, N, max_events); discretizeOCP(contact_sequence, t) fills the tables (attribute `d`)."""

    def __init__(self, T, N, max_events=None, lib=None):
        self.lib = lib or capi.default_library()
        self.T, self.N_ideal = float(T), int(N)
        self.d = capi.OCPDiscretization()

    def discretizeOCP(self, contact_sequence, t):
        self.lib.check(self.lib.L.idocp_b200_discretize_ocp(contact_sequence._h, self.T, self.N_ideal, float(t),
                                                            C.byref(self.d)))
        return bool(self.d.well_defined)

    def N(self):
        return self.d.N

    def N_impulse(self):
        return self.d.N_impulse

    def N_lift(self):
        return self.d.N_lift

    def N_all(self):
        return self.d.N + 1 + 2 * self.d.N_impulse + self.d.N_lift

    def contactPhase(self, i):
        return self.d.contact_phase[i]

    def impulseIndexAfterTimeStage(self, i):
        return self.d.impulse_index_after_time_stage[i]

    def liftIndexAfterTimeStage(self, i):
        return self.d.lift_index_after_time_stage[i]

    def timeStageBeforeImpulse(self, k):
        return self.d.time_stage_before_impulse[k]

    def timeStageBeforeLift(self, k):
        return self.d.time_stage_before_lift[k]

    def isTimeStageBeforeImpulse(self, i):
        return i < self.d.N and self.d.impulse_index_after_time_stage[i] >= 0

    def isTimeStageBeforeLift(self, i):
        return i < self.d.N and self.d.lift_index_after_time_stage[i] >= 0

    def t(self, i):
        return self.d.t[i]

    def dt(self, i):
        return self.d.dt[i]

    def t_impulse(self, k):
        return self.d.t_impulse[k]

    def t_lift(self, k):
        return self.d.t_lift[k]

    def dt_aux(self, k):
        return self.d.dt_aux[k]

    def dt_lift(self, k):
        return self.d.dt_lift[k]

    def stages(self):
        return [self.d.stages[k] for k in range(self.d.num_stages)]

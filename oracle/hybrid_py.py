"""ctypes binding of the contact-schedule oracle (oracle/hybrid_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import oracle_py

HY_MAX_CONTACTS, HY_MAX_EVENTS, HY_MAX_N = 8, 64, 1024
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class Discretization(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("N_impulse", C.c_int), ("N_lift", C.c_int), ("well_defined", C.c_int),
        ("t", C.c_double * (HY_MAX_N + 1)), ("dt", C.c_double * (HY_MAX_N + 1)),
        ("contact_phase", C.c_int * (HY_MAX_N + 1)), ("impulse_after", C.c_int * (HY_MAX_N + 1)),
        ("lift_after", C.c_int * (HY_MAX_N + 1)),
        ("before_impulse_flag", C.c_int * (HY_MAX_N + 1)), ("before_lift_flag", C.c_int * (HY_MAX_N + 1)),
        ("stage_before_impulse", C.c_int * (HY_MAX_EVENTS + 1)), ("stage_before_lift", C.c_int * (HY_MAX_EVENTS + 1)),
        ("t_impulse", C.c_double * (HY_MAX_EVENTS + 1)), ("t_lift", C.c_double * (HY_MAX_EVENTS + 1)),
        ("dt_aux", C.c_double * (HY_MAX_EVENTS + 1)), ("dt_lift", C.c_double * (HY_MAX_EVENTS + 1)),
    ]


def _lib():
    L = oracle_py.lib()
    L.oracle_cs_create.restype = C.c_void_p
    L.oracle_cs_create.argtypes = [C.c_int, C.c_int]
    L.oracle_cs_destroy.argtypes = [C.c_void_p]
    L.oracle_cs_set_uniform.argtypes = [C.c_void_p, _ip, _dp]
    L.oracle_cs_push_back.argtypes = [C.c_void_p, _ip, _dp, C.c_double]
    L.oracle_cs_pop_back.argtypes = [C.c_void_p]
    L.oracle_cs_pop_front.argtypes = [C.c_void_p]
    L.oracle_cs_update_event_time.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.oracle_cs_set_contact_points.argtypes = [C.c_void_p, C.c_int, _dp]
    L.oracle_cs_counts.argtypes = [C.c_void_p, _ip, _ip, _ip]
    L.oracle_cs_get_phase.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
    L.oracle_cs_get_impulse.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp]
    L.oracle_cs_lift_time.restype = C.c_double
    L.oracle_cs_lift_time.argtypes = [C.c_void_p, C.c_int]
    L.oracle_discretize_ocp.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.POINTER(Discretization)]
    assert L.oracle_discretization_size() == C.sizeof(Discretization)
    return L


class ContactSequence:
    def __init__(self, max_point_contacts, max_num_events):
        self.L = _lib()
        self.n = max_point_contacts
        self.h = C.c_void_p(self.L.oracle_cs_create(max_point_contacts, max_num_events))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_cs_destroy(self.h)
            self.h = None

    def _a(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=np.int32))

    def _p(self, p):
        return None if p is None else np.ascontiguousarray(np.asarray(p, dtype=np.float64))

    def set_uniform(self, active, points=None):
        a, p = self._a(active), self._p(points)
        self.L.oracle_cs_set_uniform(self.h, a.ctypes.data_as(_ip), None if p is None else p.ctypes.data_as(_dp))

    def push_back(self, active, event_time, points=None):
        a, p = self._a(active), self._p(points)
        return self.L.oracle_cs_push_back(self.h, a.ctypes.data_as(_ip), None if p is None else p.ctypes.data_as(_dp),
                                          float(event_time))

    def pop_back(self):
        self.L.oracle_cs_pop_back(self.h)

    def pop_front(self):
        self.L.oracle_cs_pop_front(self.h)

    def update_event_time(self, impulse, index, time):
        return self.L.oracle_cs_update_event_time(self.h, int(impulse), int(index), float(time))

    def set_contact_points(self, phase, points):
        p = self._p(points)
        return self.L.oracle_cs_set_contact_points(self.h, int(phase), p.ctypes.data_as(_dp))

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L.oracle_cs_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def phase(self, k):
        a, p = np.zeros(self.n, dtype=np.int32), np.zeros((self.n, 3))
        self.L.oracle_cs_get_phase(self.h, k, a.ctypes.data_as(_ip), p.ctypes.data_as(_dp))
        return a, p

    def impulse(self, k):
        a, p, t = np.zeros(self.n, dtype=np.int32), np.zeros((self.n, 3)), C.c_double()
        self.L.oracle_cs_get_impulse(self.h, k, a.ctypes.data_as(_ip), p.ctypes.data_as(_dp), C.byref(t))
        return a, p, t.value

    def lift_time(self, k):
        return self.L.oracle_cs_lift_time(self.h, k)

    def discretize(self, T, N, t):
        d = Discretization()
        assert self.L.oracle_discretize_ocp(self.h, float(T), int(N), float(t), C.byref(d)) == 0
        return d

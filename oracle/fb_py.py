"""ctypes binding of the floating-base part of the CPU oracle (oracle/fb_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import oracle_py

NV, NQ, NU = 18, 19, 12
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _a(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def lib():
    L = oracle_py.lib()
    if not getattr(L, "_fb_ready", False):
        for name in ("oracle_fb_integrate", "oracle_fb_rnea"):
            getattr(L, name)
        L.oracle_fb_integrate.argtypes = [_dp, _dp, C.c_double, _dp]
        L.oracle_fb_rnea.argtypes = [_dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]
        L.oracle_fb_contact.argtypes = [_dp, _dp, _dp, C.c_int, C.c_double] + [_dp] * 12
        L.oracle_fb_mjtjinv.argtypes = [_dp, _dp, C.c_int, _dp]
        L._fb_ready = True
    return L


def integrate(q, v, alpha=1.0):
    out = np.zeros(NQ)
    lib().oracle_fb_integrate(_p(_a(q, NQ)), _p(_a(v, NV)), float(alpha), _p(out))
    return out


def subtract(q_plus, q_minus):
    out = np.zeros(NV)
    lib().oracle_fb_subtract(_p(_a(q_plus, NQ)), _p(_a(q_minus, NQ)), _p(out))
    return out


def dsubtract(q_plus, q_minus):
    jp, jm = np.zeros((6, 6)), np.zeros((6, 6))
    lib().oracle_fb_dsubtract(_p(_a(q_plus, NQ)), _p(_a(q_minus, NQ)), _p(jp), _p(jm))
    return jp, jm


def dsubtract_inverse(J):
    out = np.zeros((6, 6))
    lib().oracle_fb_dsubtract_inverse(_p(_a(J, 36)), _p(out))
    return out


def dintegrate(v6):
    jq, jv = np.zeros((6, 6)), np.zeros((6, 6))
    lib().oracle_fb_dintegrate(_p(_a(v6, 6)), _p(jq), _p(jv))
    return jq, jv


def exp6(nu):
    R, p = np.zeros((3, 3)), np.zeros(3)
    lib().oracle_fb_exp6(_p(_a(nu, 6)), _p(R), _p(p))
    return R, p


def log6(R, p):
    out = np.zeros(6)
    lib().oracle_fb_log6(_p(_a(R, 9)), _p(_a(p, 3)), _p(out))
    return out


def rnea(q, v, a, f=None, gravity=9.81, derivatives=False):
    tau = np.zeros(NV)
    f12 = _a(f, 12) if f is not None else None
    if not derivatives:
        lib().oracle_fb_rnea(_p(_a(q, NQ)), _p(_a(v, NV)), _p(_a(a, NV)), _p(f12), gravity, _p(tau), None, None, None)
        return tau
    dq, dv, M = np.zeros((NV, NV)), np.zeros((NV, NV)), np.zeros((NV, NV))
    lib().oracle_fb_rnea(_p(_a(q, NQ)), _p(_a(v, NV)), _p(_a(a, NV)), _p(f12), gravity, _p(tau), _p(dq), _p(dv), _p(M))
    return tau, dq, dv, M


def contact(q, v, a, i, time_step, contact_point):
    o = dict(P=np.zeros(3), vF=np.zeros(6), aF=np.zeros(6), J=np.zeros((6, NV)), v_dq=np.zeros((6, NV)),
             a_dq=np.zeros((6, NV)), a_dv=np.zeros((6, NV)), C=np.zeros(3), dCdq=np.zeros((3, NV)),
             dCdv=np.zeros((3, NV)), dCda=np.zeros((3, NV)))
    lib().oracle_fb_contact(_p(_a(q, NQ)), _p(_a(v, NV)), _p(_a(a, NV)), int(i), float(time_step), _p(_a(contact_point, 3)),
                            *[_p(o[k]) for k in ("P", "vF", "aF", "J", "v_dq", "a_dq", "a_dv", "C", "dCdq", "dCdv", "dCda")])
    return o


def mjtjinv(M, J):
    J = _a(J).reshape(-1, NV) if np.size(J) else np.zeros((0, NV))
    dimf = J.shape[0]
    out = np.zeros((NV + dimf, NV + dimf))
    info = lib().oracle_fb_mjtjinv(_p(_a(M, NV * NV)), _p(J), dimf, _p(out))
    return out, info


# ------------------------------------------------------------------------------------------------
# OCPSolver oracle (oracle/fb_ocp.c)
# ------------------------------------------------------------------------------------------------
K_GRID, K_IMPULSE, K_AUX, K_LIFT, K_TERMINAL = range(5)
NCOMP = 8


class FbProblem(C.Structure):
    _fields_ = [
        ("T", C.c_double), ("N", C.c_int), ("max_num_impulse", C.c_int),
        ("q_weight", C.c_double * NV), ("v_weight", C.c_double * NV), ("a_weight", C.c_double * NV),
        ("qf_weight", C.c_double * NV), ("vf_weight", C.c_double * NV), ("qi_weight", C.c_double * NV),
        ("vi_weight", C.c_double * NV), ("dvi_weight", C.c_double * NV),
        ("f_weight", C.c_double * 12), ("f_ref", C.c_double * 12), ("fi_weight", C.c_double * 12), ("fi_ref", C.c_double * 12),
        ("q_min", C.c_double * NU), ("q_max", C.c_double * NU), ("v_max", C.c_double * NU), ("u_max", C.c_double * NU),
        ("mu", C.c_double), ("barrier", C.c_double), ("fraction_rate", C.c_double),
        ("enable", C.c_int * NCOMP),
        ("cone_nonlinear", C.c_int * 2), ("enable_acc", C.c_int * 2), ("a_min", C.c_double * NU), ("a_max", C.c_double * NU),
        ("enable_distance", C.c_int),
    ]

    def set(self, name, values):
        arr = getattr(self, name)
        vals = np.asarray(values, dtype=float).ravel()
        assert len(arr) == vals.size, (name, len(arr), vals.size)
        for i, x in enumerate(vals):
            arr[i] = x


_SIZES = dict(ls_cost=1, ls_viol=1, q=19, f=12, mu=12, nu_passive=6, xi=12, u=12, du=12, daf=30, dbetamu=30, dnu_passive=6, dxi=12, lf=12, lu=12,
              lu_passive=6, P=12, IDC=30, Qxx=36 * 36, Qxu=36 * 18, Quu=18 * 18, Qff=144, Fvq=324, Fvv=324, Fvu=216, Fqq6=36, Fqv6=36,
              MJtJinv=900, MJ_dIDC=30 * 36, MJ_IDC=30, dIDCdqv=30 * 36, dCda=12 * 18, Mm=324, K=12 * 36, k=12, Pqq=324, Pqv=324,
              Pvv=324, Phix=12 * 36, Phia=12 * 18, Phiu=144, cM=12 * 36, cm=12, Fqq_prev_inv=36, Fqq_inv=36, laf=30,
              Qafqv=30 * 36, Qafu=30 * 18, kkt=1, active=4, cdJ=72, cdz=4, slack=6 * 12 + 40 + 24 + 4, dual=6 * 12 + 40 + 24 + 4)


class FbOCP:
    def __init__(self, problem, contact_sequence):
        self.L = lib()
        L = self.L
        L.oracle_fb_ocp_create.restype = C.c_void_p
        L.oracle_fb_ocp_create.argtypes = [C.POINTER(FbProblem), C.c_void_p]
        L.oracle_fb_ocp_destroy.argtypes = [C.c_void_p]
        L.oracle_fb_ocp_set_solution.argtypes = [C.c_void_p, C.c_char_p, _dp]
        L.oracle_fb_ocp_set_reference.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.oracle_fb_ocp_discretize.argtypes = [C.c_void_p, C.c_double]
        L.oracle_fb_ocp_chain.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 2 + [_dp] * 2 + [C.POINTER(C.c_int)] * 2
        L.oracle_fb_ocp_init_constraints.argtypes = [C.c_void_p, C.c_double]
        L.oracle_fb_ocp_update_solution.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        L.oracle_fb_ocp_update_solution_ls.argtypes = [C.c_void_p, C.c_double, _dp, _dp, C.c_int]
        L.oracle_fb_ocp_clear_line_search_filter.argtypes = [C.c_void_p]
        L.oracle_fb_ocp_filter_size.argtypes = [C.c_void_p]
        L.oracle_fb_ocp_cost_and_violation.argtypes = [C.c_void_p, C.c_double, _dp]
        L.oracle_fb_ocp_compute_kkt_residual.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        L.oracle_fb_ocp_kkt_error.restype = C.c_double
        L.oracle_fb_ocp_kkt_error.argtypes = [C.c_void_p]
        L.oracle_fb_ocp_get_step_sizes.argtypes = [C.c_void_p, _dp]
        L.oracle_fb_ocp_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _dp]
        L.oracle_fb_ocp_set.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _dp]
        L.oracle_fb_ocp_set_threads.argtypes = [C.c_void_p, C.c_int]
        assert L.oracle_fb_problem_size() == C.sizeof(FbProblem)
        self.problem = problem
        self.cs = contact_sequence
        self.h = C.c_void_p(L.oracle_fb_ocp_create(C.byref(problem), contact_sequence.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_fb_ocp_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self.L.oracle_fb_ocp_set_threads(self.h, int(n))

    def set_solution(self, name, value):
        assert self.L.oracle_fb_ocp_set_solution(self.h, name.encode(), _p(_a(value))) == 0

    def set_reference(self, kind, index, q_ref, v_ref):
        self.L.oracle_fb_ocp_set_reference(self.h, int(kind), int(index), _p(_a(q_ref, NQ)), _p(_a(v_ref, NV)))

    def discretize(self, t):
        return self.L.oracle_fb_ocp_discretize(self.h, float(t))

    def chain(self):
        n = 2048
        kind, index = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t, dt = np.zeros(n), np.zeros(n)
        dimf, dimi = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ip = C.POINTER(C.c_int)
        m = self.L.oracle_fb_ocp_chain(self.h, kind.ctypes.data_as(ip), index.ctypes.data_as(ip), _p(t), _p(dt),
                                       dimf.ctypes.data_as(ip), dimi.ctypes.data_as(ip))
        return [dict(kind=int(kind[e]), index=int(index[e]), t=float(t[e]), dt=float(dt[e]), dimf=int(dimf[e]), dimi=int(dimi[e]))
                for e in range(m)]

    def init_constraints(self, t):
        return self.L.oracle_fb_ocp_init_constraints(self.h, float(t))

    def update_solution(self, t, q, v, line_search=False):
        return self.L.oracle_fb_ocp_update_solution_ls(self.h, float(t), _p(_a(q, NQ)), _p(_a(v, NV)), int(line_search))

    def clear_line_search_filter(self):
        self.L.oracle_fb_ocp_clear_line_search_filter(self.h)

    def filter_size(self):
        return self.L.oracle_fb_ocp_filter_size(self.h)

    def cost_and_violation(self, alpha):
        out = np.zeros(2)
        self.L.oracle_fb_ocp_cost_and_violation(self.h, float(alpha), _p(out))
        return out

    def compute_kkt_residual(self, t, q, v):
        return self.L.oracle_fb_ocp_compute_kkt_residual(self.h, float(t), _p(_a(q, NQ)), _p(_a(v, NV)))

    def kkt_error(self):
        return self.L.oracle_fb_ocp_kkt_error(self.h)

    def step_sizes(self):
        out = np.zeros(2)
        self.L.oracle_fb_ocp_get_step_sizes(self.h, _p(out))
        return out

    def get(self, e, name):
        out = np.zeros(_SIZES.get(name, NV))
        n = self.L.oracle_fb_ocp_get(self.h, int(e), name.encode(), _p(out))
        assert n == out.size, (name, n, out.size)
        return out

    def set(self, e, name, value):
        assert self.L.oracle_fb_ocp_set(self.h, int(e), name.encode(), _p(_a(value))) == 0


def batch_update_solution(solvers, t, q, v, line_search=False, nthreads=1):
    """OpenMP over instances (one FbOCP per instance), every instance single-threaded."""
    L = lib()
    L.oracle_fb_ocp_batch_update_solution.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, _dp, _dp, C.c_int, C.c_int]
    arr = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    L.oracle_fb_ocp_batch_update_solution(arr, len(solvers), float(t), _p(_a(q, len(solvers) * NQ)), _p(_a(v, len(solvers) * NV)),
                                          int(line_search), int(nthreads))


def _harr(solvers):
    return (C.c_void_p * len(solvers))(*[s.h for s in solvers])


def batch_kkt(solvers, t, q, v, nthreads=1):
    """computeKKTResidual + KKTError of every instance (OpenMP over instances)."""
    L = lib()
    L.oracle_fb_ocp_batch_kkt.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, _dp, _dp, _dp, C.c_int]
    out = np.zeros(len(solvers))
    L.oracle_fb_ocp_batch_kkt(_harr(solvers), len(solvers), float(t), _p(_a(q, len(solvers) * NQ)), _p(_a(v, len(solvers) * NV)),
                              _p(out), int(nthreads))
    return out


def batch_get(solvers, e, name):
    """(batch, size) of one field of chain element e of every instance."""
    L = lib()
    L.oracle_fb_ocp_batch_get.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_char_p, _dp, C.c_int]
    size = _SIZES.get(name, NV)
    out = np.zeros((len(solvers), size))
    n = L.oracle_fb_ocp_batch_get(_harr(solvers), len(solvers), int(e), name.encode(), _p(out), size)
    assert n == size, (name, n, size)
    return out


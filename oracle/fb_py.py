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

"""Independent numpy formulation of the ANYmal rigid-body model -- TEST INFRASTRUCTURE ONLY.

Classic Featherstone body-frame RNEA on the kinematic tree (free-flyer + 4x3 revolute joints) with external
forces, written without reference to oracle/fb_robot.h (which works in the world frame), plus SE(3) helpers
through scipy's expm/logm.  Used to validate the oracle at the pinocchio boundary."""
import json
import os

import numpy as np
from scipy.linalg import expm, logm

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_model():
    with open(os.path.join(_HERE, "..", "tests", "golden", "model_anymal.json")) as f:
        return json.load(f)


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_axis(ax, th):
    c, s = np.cos(th), np.sin(th)
    if ax == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if ax == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def hat6(nu):
    T = np.zeros((4, 4))
    T[:3, :3] = skew(nu[3:])
    T[:3, 3] = nu[:3]
    return T


def exp6(nu):
    T = expm(hat6(np.asarray(nu, float)))
    return T[:3, :3], T[:3, 3]


def log6(R, p):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = p
    L = np.real(logm(T))
    return np.array([L[0, 3], L[1, 3], L[2, 3], L[2, 1], L[0, 2], L[1, 0]])


def integrate(q, v, alpha=1.0):
    q = np.asarray(q, float)
    R, p = exp6(alpha * np.asarray(v[:6], float))
    R0 = quat_to_R(q[3:7])
    out = q.copy()
    out[:3] = q[:3] + R0 @ p
    R1 = R0 @ R
    from scipy.spatial.transform import Rotation
    qq = Rotation.from_matrix(R1).as_quat()
    if qq @ q[3:7] < 0:
        qq = -qq
    out[3:7] = qq
    out[7:] = q[7:] + alpha * np.asarray(v[6:], float)
    return out


def rnea_body(m, q, v, a, f=None, gravity=9.81):
    """Body-frame recursive Newton-Euler; spatial vectors [linear; angular] in each joint frame."""
    q, v, a = (np.asarray(x, float) for x in (q, v, a))
    f = np.zeros((4, 3)) if f is None else np.asarray(f, float).reshape(4, 3)
    nb = 13
    Rw = [None] * nb      # world rotation of body frame
    vl, va, al, aa = ([None] * nb for _ in range(4))
    R0 = quat_to_R(q[3:7])
    Rw[0] = R0
    vl[0], va[0] = v[:3].copy(), v[3:6].copy()
    # classical (non-spatial) bookkeeping in body frames: use spatial accel with gravity trick
    # spatial acceleration of base in its own frame: a_lin + (gravity), note spatial accel = dv/dt in moving coords
    al[0] = a[:3] + R0.T @ np.array([0, 0, gravity])
    aa[0] = a[3:6].copy()
    Rpar = [None] * nb
    ppar = [None] * nb
    for j in range(12):
        b = 1 + j
        pb = m["parent"][j] + 1
        Rl = rot_axis(m["axis"][j], q[7 + j])
        pl = np.array(m["p"][j])
        Rpar[b], ppar[b] = Rl, pl
        Rw[b] = Rw[pb] @ Rl
        e = np.eye(3)[m["axis"][j]]
        # motion transform parent -> child: X = [R^T, -R^T [p]x; 0, R^T]
        vpl = Rl.T @ (vl[pb] + np.cross(va[pb], pl))
        vpa = Rl.T @ va[pb]
        vl[b] = vpl
        va[b] = vpa + e * v[6 + j]
        apl = Rl.T @ (al[pb] + np.cross(aa[pb], pl))
        apa = Rl.T @ aa[pb]
        # a_child = X a_parent + S qdd + v_child x (S qd)
        sl, sa = np.zeros(3), e * v[6 + j]
        cl = np.cross(va[b], sl) + np.cross(vl[b], sa)
        ca = np.cross(va[b], sa)
        al[b] = apl + cl
        aa[b] = apa + e * a[6 + j] + ca
    fl, fa = [None] * nb, [None] * nb
    for b in range(nb):
        mass = m["mass"][b]
        c = np.array(m["com"][b])
        Ic = np.array(m["inertia"][b])
        Io = Ic - mass * skew(c) @ skew(c)

        def Y(l, w):
            return mass * l - mass * np.cross(c, w), Io @ w + mass * np.cross(c, l)
        hl, ha = Y(vl[b], va[b])
        yl, ya = Y(al[b], aa[b])
        fl[b] = yl + np.cross(va[b], hl)
        fa[b] = ya + np.cross(va[b], ha) + np.cross(vl[b], hl)
    for i in range(4):
        b = 1 + m["contact_parent"][i]
        pc = np.array(m["contact_p"][i])
        fl[b] = fl[b] - f[i]
        fa[b] = fa[b] - np.cross(pc, f[i])
    tau = np.zeros(18)
    for b in range(nb - 1, 0, -1):
        j = b - 1
        pb = m["parent"][j] + 1
        e = np.eye(3)[m["axis"][j]]
        tau[6 + j] = e @ fa[b]
        Rl, pl = Rpar[b], ppar[b]
        fpl = Rl @ fl[b]
        fl[pb] = fl[pb] + fpl
        fa[pb] = fa[pb] + Rl @ fa[b] + np.cross(pl, fpl)
    tau[:3] = fl[0]
    tau[3:6] = fa[0]
    return tau


def frame_state(m, q, v, a, i):
    """World position, LOCAL velocity and LOCAL classical acceleration of contact frame i (direct recursion)."""
    q, v, a = (np.asarray(x, float) for x in (q, v, a))
    R = quat_to_R(q[3:7])
    p = q[:3].copy()
    vl, va = v[:3].copy(), v[3:6].copy()     # body-frame spatial velocity
    al, aa = a[:3].copy(), a[3:6].copy()     # body-frame spatial acceleration (no gravity)
    leg = i
    for jj in range(3):
        j = 3 * leg + jj
        Rl = rot_axis(m["axis"][j], q[7 + j])
        pl = np.array(m["p"][j])
        e = np.eye(3)[m["axis"][j]]
        p = p + R @ pl
        R = R @ Rl
        nvl = Rl.T @ (vl + np.cross(va, pl))
        nva = Rl.T @ va + e * v[6 + j]
        sa = e * v[6 + j]
        nal = Rl.T @ (al + np.cross(aa, pl)) + np.cross(nvl, sa)
        naa = Rl.T @ aa + e * a[6 + j] + np.cross(nva, sa)
        vl, va, al, aa = nvl, nva, nal, naa
    pc = np.array(m["contact_p"][i])
    P = p + R @ pc
    fvl = vl + np.cross(va, pc)
    fal = al + np.cross(aa, pc)
    acl = fal + np.cross(va, fvl)
    return P, R, np.concatenate([fvl, va]), np.concatenate([fal, aa]), acl

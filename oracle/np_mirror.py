"""NumPy mirror of the rigid-body part of the oracle (TEST INFRASTRUCTURE ONLY).

Independent restatement used to cross-check oracle/idocp_oracle.c:
  * rnea_body(): Featherstone's recursive Newton-Euler in BODY frames, the formulation
    pinocchio::rnea uses (call site: reference include/idocp/robot/robot.hxx:444-460);
  * rnea_derivatives_world(): analytical derivatives in the WORLD frame (Carpentier & Mansard,
    RSS 2018 -- the algorithm behind pinocchio::computeRNEADerivatives, call site robot.hxx:466-500);
  * rnea_derivatives_fd(): central finite differences of rnea_body (independent check).
Spatial vectors are [linear; angular] as in pinocchio.  Only tests/ may import this module.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_model(name="iiwa14"):
    with open(os.path.join(_HERE, "..", "tests", "golden", "model_%s.json" % name)) as f:
        m = json.load(f)
    for k in ("R", "p", "mass", "com", "inertia", "q_min", "q_max", "v_max", "effort"):
        m[k] = np.array(m[k], dtype=float)
    return m


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0.0]])


def rotz(q):
    c, s = np.cos(q), np.sin(q)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def crm(v):
    """motion cross-product matrix: crm(v) @ m = v x m ([lin; ang] ordering)."""
    X = np.zeros((6, 6))
    X[:3, :3] = skew(v[3:])
    X[:3, 3:] = skew(v[:3])
    X[3:, 3:] = skew(v[3:])
    return X


def crf(v):
    """force cross-product matrix: crf(v) @ f = v x* f."""
    return -crm(v).T


def spatial_inertia(m, c, Ic):
    """6x6 inertia about the frame origin from (mass, com, inertia about com)."""
    I = np.zeros((6, 6))
    C = skew(c)
    I[:3, :3] = m * np.eye(3)
    I[:3, 3:] = -m * C
    I[3:, :3] = m * C
    I[3:, 3:] = Ic - m * C @ C
    return I


def rnea_body(model, q, v, a, gravity=9.81):
    n = len(q)
    vs, as_, fs, Xs = [], [], [], []
    S = np.array([0, 0, 0, 0, 0, 1.0])
    for i in range(n):
        R = model["R"][i] @ rotz(q[i])          # child -> parent rotation
        p = model["p"][i]
        # actInv: parent-frame motion expressed in the child frame
        def act_inv(m_):
            lin = R.T @ (m_[:3] - np.cross(p, m_[3:]))
            return np.concatenate([lin, R.T @ m_[3:]])
        vp = vs[i - 1] if i > 0 else np.zeros(6)
        ap = as_[i - 1] if i > 0 else np.array([0, 0, gravity, 0, 0, 0.0])
        vi = S * v[i] + act_inv(vp)
        ai = S * a[i] + crm(vi) @ (S * v[i]) + act_inv(ap)
        I = spatial_inertia(model["mass"][i], model["com"][i], model["inertia"][i])
        fi = I @ ai + crf(vi) @ (I @ vi)
        vs.append(vi); as_.append(ai); fs.append(fi); Xs.append((R, p))
    tau = np.zeros(n)
    for i in reversed(range(n)):
        tau[i] = S @ fs[i]
        if i > 0:
            R, p = Xs[i]
            lin = R @ fs[i][:3]
            ang = R @ fs[i][3:] + np.cross(p, lin)
            fs[i - 1] = fs[i - 1] + np.concatenate([lin, ang])
    return tau


def world_quantities(model, q, v, a, gravity=9.81):
    """Forward sweep in the world frame: S, dS, v, acc, B, I, D, f per joint."""
    n = len(q)
    R = np.eye(3); p = np.zeros(3)
    vw = np.zeros(6); aw = np.array([0, 0, gravity, 0, 0, 0.0])
    out = []
    for i in range(n):
        p = p + R @ model["p"][i]
        R = R @ model["R"][i] @ rotz(q[i])
        z = R[:, 2]
        S = np.concatenate([np.cross(p, z), z])
        vw = vw + S * v[i]
        dS = crm(vw) @ S
        aw = aw + S * a[i] + dS * v[i]
        B = crm(aw) @ S + crm(vw) @ dS
        cw = R @ model["com"][i] + p
        I = spatial_inertia(model["mass"][i], cw, R @ model["inertia"][i] @ R.T)
        h = I @ vw
        f = I @ aw + crf(vw) @ h
        # D m = I (m x v) + m x* h + v x* (I m)
        Hx = np.zeros((6, 6))
        Hx[:3, 3:] = -skew(h[:3])
        Hx[3:, :3] = -skew(h[:3])
        Hx[3:, 3:] = -skew(h[3:])
        D = -I @ crm(vw) + Hx + crf(vw) @ I
        out.append(dict(S=S, dS=dS, v=vw.copy(), a=aw.copy(), B=B, I=I, D=D, f=f))
    return out


def rnea_derivatives_world(model, q, v, a, gravity=9.81):
    n = len(q)
    w = world_quantities(model, q, v, a, gravity)
    IC = np.zeros((6, 6)); DC = np.zeros((6, 6)); F = np.zeros(6)
    U = [None] * n; W = [None] * n; G = [None] * n; H = [None] * n
    tau = np.zeros(n)
    for i in reversed(range(n)):
        IC = IC + w[i]["I"]; DC = DC + w[i]["D"]; F = F + w[i]["f"]
        S, dS, B = w[i]["S"], w[i]["dS"], w[i]["B"]
        tau[i] = S @ F
        U[i] = IC @ S
        W[i] = DC.T @ S
        G[i] = crf(S) @ F + IC @ B + DC @ dS
        H[i] = DC @ S + 2.0 * (IC @ dS)
    dq = np.zeros((n, n)); dv = np.zeros((n, n)); M = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i <= j:
                dq[i, j] = w[i]["S"] @ G[j]
                dv[i, j] = w[i]["S"] @ H[j]
                M[i, j] = w[i]["S"] @ U[j]
                M[j, i] = M[i, j]
            else:
                dq[i, j] = U[i] @ w[j]["B"] + W[i] @ w[j]["dS"]
                dv[i, j] = W[i] @ w[j]["S"] + 2.0 * (U[i] @ w[j]["dS"])
    return tau, dq, dv, M


def rnea_derivatives_fd(model, q, v, a, eps=1e-6):
    n = len(q)
    dq = np.zeros((n, n)); dv = np.zeros((n, n)); da = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n); e[j] = eps
        dq[:, j] = (rnea_body(model, q + e, v, a) - rnea_body(model, q - e, v, a)) / (2 * eps)
        dv[:, j] = (rnea_body(model, q, v + e, a) - rnea_body(model, q, v - e, a)) / (2 * eps)
        da[:, j] = (rnea_body(model, q, v, a + e) - rnea_body(model, q, v, a - e)) / (2 * eps)
    return dq, dv, da


def rnea_derivatives_compact(model, q, v, a, gravity=9.81):
    """Same algorithm as rnea_derivatives_world with the structure of the composite matrices
    exploited (3-vector algebra only).  This is the operation order oracle/idocp_oracle.c and the
    CUDA kernel follow; kept here so the three can be compared term by term.
      I^C  -> (m, mc, Ibar)            10 numbers
      D^C  -> (hl, ha, Sym)            12 numbers:  D m = (-2 hl x m_w ; Sym m_w - ha x m_w)
    """
    n = len(q)
    cr = np.cross
    R = np.eye(3); p = np.zeros(3)
    vl = np.zeros(3); vw = np.zeros(3); al = np.array([0, 0, gravity]); aw = np.zeros(3)
    J = []
    for i in range(n):
        p = p + R @ model["p"][i]
        R = R @ model["R"][i] @ rotz(q[i])
        z = R[:, 2]
        Sl = cr(p, z); Sw = z
        vw = vw + Sw * v[i]; vl = vl + Sl * v[i]
        dSl = cr(vw, Sl) + cr(vl, Sw); dSw = cr(vw, Sw)
        aw = aw + Sw * a[i] + dSw * v[i]; al = al + Sl * a[i] + dSl * v[i]
        Bl = cr(aw, Sl) + cr(al, Sw) + cr(vw, dSl) + cr(vl, dSw)
        Bw = cr(aw, Sw) + cr(vw, dSw)
        m = model["mass"][i]
        cw = R @ model["com"][i] + p
        mc = m * cw
        Ibar = R @ model["inertia"][i] @ R.T + m * ((cw @ cw) * np.eye(3) - np.outer(cw, cw))
        hl = m * vl + cr(vw, mc); ha = cr(mc, vl) + Ibar @ vw
        fl = m * al + cr(aw, mc) + cr(vw, hl)
        fa = cr(mc, al) + Ibar @ aw + cr(vw, ha) + cr(vl, hl)
        wI = skew(vw) @ Ibar
        Sym = -(np.outer(vl, mc) + np.outer(mc, vl)) + 2 * (mc @ vl) * np.eye(3) + wI + wI.T
        J.append(dict(Sl=Sl, Sw=Sw, dSl=dSl, dSw=dSw, Bl=Bl, Bw=Bw, m=m, mc=mc, Ibar=Ibar,
                      hl=hl, ha=ha, Sym=Sym, fl=fl, fa=fa))
    mC = 0.0; mcC = np.zeros(3); IC = np.zeros((3, 3)); hlC = np.zeros(3); haC = np.zeros(3)
    SymC = np.zeros((3, 3)); Fl = np.zeros(3); Fa = np.zeros(3)
    tau = np.zeros(n)
    for i in reversed(range(n)):
        j = J[i]
        mC += j["m"]; mcC = mcC + j["mc"]; IC = IC + j["Ibar"]; hlC = hlC + j["hl"]; haC = haC + j["ha"]
        SymC = SymC + j["Sym"]; Fl = Fl + j["fl"]; Fa = Fa + j["fa"]
        Sl, Sw, dSl, dSw, Bl, Bw = j["Sl"], j["Sw"], j["dSl"], j["dSw"], j["Bl"], j["Bw"]
        tau[i] = Sl @ Fl + Sw @ Fa
        j["Ul"] = mC * Sl + cr(Sw, mcC); j["Uw"] = cr(mcC, Sl) + IC @ Sw
        j["Ww"] = 2 * cr(hlC, Sl) + SymC @ Sw + cr(haC, Sw)
        j["Gl"] = cr(Sw, Fl) + mC * Bl + cr(Bw, mcC) - 2 * cr(hlC, dSw)
        j["Gw"] = cr(Sw, Fa) + cr(Sl, Fl) + cr(mcC, Bl) + IC @ Bw + SymC @ dSw - cr(haC, dSw)
        j["Hl"] = -2 * cr(hlC, Sw) + 2 * (mC * dSl + cr(dSw, mcC))
        j["Hw"] = SymC @ Sw - cr(haC, Sw) + 2 * (cr(mcC, dSl) + IC @ dSw)
    dq = np.zeros((n, n)); dv = np.zeros((n, n)); M = np.zeros((n, n))
    for i in range(n):
        for k in range(n):
            a_, b_ = J[i], J[k]
            if i <= k:
                dq[i, k] = a_["Sl"] @ b_["Gl"] + a_["Sw"] @ b_["Gw"]
                dv[i, k] = a_["Sl"] @ b_["Hl"] + a_["Sw"] @ b_["Hw"]
                M[i, k] = a_["Sl"] @ b_["Ul"] + a_["Sw"] @ b_["Uw"]
                M[k, i] = M[i, k]
            else:
                dq[i, k] = a_["Ul"] @ b_["Bl"] + a_["Uw"] @ b_["Bw"] + a_["Ww"] @ b_["dSw"]
                dv[i, k] = a_["Ww"] @ b_["Sw"] + 2 * (a_["Ul"] @ b_["dSl"] + a_["Uw"] @ b_["dSw"])
    return tau, dq, dv, M

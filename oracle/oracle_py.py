"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package idocp_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NV = 7
NC = 6
_dp = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("T", C.c_double),
        ("q_ref", C.c_double * NV), ("v_ref", C.c_double * NV), ("u_ref", C.c_double * NV),
        ("q_weight", C.c_double * NV), ("v_weight", C.c_double * NV), ("a_weight", C.c_double * NV),
        ("u_weight", C.c_double * NV), ("qf_weight", C.c_double * NV), ("vf_weight", C.c_double * NV),
        ("q_min", C.c_double * NV), ("q_max", C.c_double * NV), ("v_max", C.c_double * NV),
        ("u_max", C.c_double * NV),
        ("barrier", C.c_double), ("fraction_rate", C.c_double),
        ("task_enabled", C.c_int),
        ("task_q_weight", C.c_double * 6), ("task_qf_weight", C.c_double * 6),
        ("task_center", C.c_double * 3), ("task_radius", C.c_double),
        ("task_t0", C.c_double), ("task_tf", C.c_double),
        ("task_rot_ref", C.c_double * 9),
        ("enable_acc", C.c_int * 2), ("a_min", C.c_double * NV), ("a_max", C.c_double * NV),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_unocp_create.restype = C.c_void_p
        L.oracle_unocp_kkt_error.restype = C.c_double
        L.oracle_splitmix_uniform.restype = C.c_double
        L.oracle_splitmix_uniform.argtypes = [C.c_ulonglong, C.c_ulonglong]
        if hasattr(L, "oracle_unparnmpc_create"):
            L.oracle_unparnmpc_create.restype = C.c_void_p
            L.oracle_unparnmpc_kkt_error.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def default_problem():
    p = Problem()
    lib().oracle_problem_default(C.byref(p))
    return p


def benchmark_problem(N=20, T=1.0):
    """examples/iiwa14/unocp_benchmark.cpp:22-46."""
    p = default_problem()
    p.N, p.T = N, T
    for i in range(NV):
        p.u_max[i] = 200.0
        p.q_ref[i] = -5.0
        p.v_ref[i] = -9.0
        p.q_weight[i] = 10.0
        p.qf_weight[i] = 10.0
        p.v_weight[i] = 0.1
        p.vf_weight[i] = 0.1
        p.a_weight[i] = 0.01
        p.u_weight[i] = 0.0
    return p


def config_space_problem():
    """examples/iiwa14/config_space_ocp.cpp:26-61."""
    p = default_problem()
    p.N, p.T = 60, 3.0
    qref = [0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]
    for i in range(NV):
        p.u_max[i] = 50.0
        p.v_max[i] = np.pi / 2
        p.q_ref[i] = qref[i]
        p.q_weight[i] = 10.0
        p.qf_weight[i] = 10.0
        p.v_weight[i] = 0.01
        p.vf_weight[i] = 0.01
        p.a_weight[i] = 0.01
    return p


def task_space_problem(N=120, T=6.0):
    """examples/iiwa14/task_space_ocp.cpp:55-84 (cost / limits; the 6D reference is sampled by task_space_ref)."""
    p = default_problem()
    p.N, p.T = N, T
    for i in range(NV):
        p.u_max[i] = 50.0
        p.v_max[i] = np.pi / 2
        p.q_weight[i] = 0.0
        p.qf_weight[i] = 0.0
        p.v_weight[i] = 0.01
        p.vf_weight[i] = 0.01
        p.a_weight[i] = 0.01
    p.task_enabled = 1
    for k in range(6):
        p.task_q_weight[k] = 1000.0
        p.task_qf_weight[k] = 1000.0
    return p


def task_space_ref(t):
    """TimeVaryingTaskSpace6DRef::compute_q_6d_ref of examples/iiwa14/task_space_ocp.cpp:21-46 -> [R row-major, p]."""
    rotm = [0.0, 0.0, 1.0, 0.0, 1.0, 0.0, -1.0, 0.0, 0.0]
    pos = [0.546, 0.0 + 0.1 * np.sin(np.pi * t), 0.76 + 0.1 * np.cos(np.pi * t)]
    return np.array(rotm + pos)


def task_ref_table(ref_fn, t, T, N, kind="unocp"):
    """Rows = the SE3 reference at the time each stage index is linearised at (see oracle_unocp_set_task_ref)."""
    dt = T / N
    if kind == "unocp":
        times = [t + i * dt for i in range(N)] + [t + T]
    else:
        times = [t + (i + 1) * dt for i in range(N - 1)] + [t + T, t + N * dt]
    return np.ascontiguousarray(np.array([ref_fn(x) for x in times]))


def rnea(q, v, a):
    q, v, a = _vec(q), _vec(v), _vec(a)
    tau = np.zeros(NV)
    lib().oracle_rnea(_p(q), _p(v), _p(a), _p(tau))
    return tau


def rnea_derivatives(q, v, a):
    q, v, a = _vec(q), _vec(v), _vec(a)
    dq, dv, da = np.zeros((NV, NV)), np.zeros((NV, NV)), np.zeros((NV, NV))
    lib().oracle_rnea_derivatives(_p(q), _p(v), _p(a), _p(dq), _p(dv), _p(da))
    return dq.T.copy(), dv.T.copy(), da.T.copy()   # column-major -> numpy [row, col]


def splitmix_uniform(seed, index):
    return lib().oracle_splitmix_uniform(seed, index)


class _SolverBase:
    _prefix = None

    def __init__(self, problem):
        self.L = lib()
        self.N = problem.N
        self.problem = problem
        self.h = C.c_void_p(getattr(self.L, self._prefix + "create")(C.byref(problem)))
        if not self.h:
            raise ValueError("invalid problem")

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.L, self._prefix + "destroy")(self.h)
            self.h = None

    def _f(self, name):
        return getattr(self.L, self._prefix + name)

    def set_solution(self, name, value):
        value = _vec(value)
        if self._f("set_solution")(self.h, name.encode(), _p(value)) != 0:
            raise ValueError(name)

    def init_constraints(self):
        self._f("init_constraints")(self.h)

    def set_task_ref(self, table):
        table = _vec(table)
        assert table.shape == (self.N + 1, 12)
        self._f("set_task_ref")(self.h, _p(table))

    def update_solution(self, t, q, v, line_search=False):
        q, v = _vec(q), _vec(v)
        self._f("update_solution")(self.h, C.c_double(t), _p(q), _p(v), int(line_search))

    def compute_kkt_residual(self, t, q, v):
        q, v = _vec(q), _vec(v)
        self._f("compute_kkt_residual")(self.h, C.c_double(t), _p(q), _p(v))

    def kkt_error(self):
        return self._f("kkt_error")(self.h)

    def _rows(self, name):
        raise NotImplementedError

    def get_solution(self, name):
        out = np.zeros((self.N + 1, NV))
        n = self._f("get_solution")(self.h, name.encode(), _p(out))
        if n < 0:
            raise ValueError(name)
        return out[:n].copy()

    def get_direction(self, name):
        out = np.zeros((self.N + 1, NV))
        n = self._f("get_direction")(self.h, name.encode(), _p(out))
        if n < 0:
            raise ValueError(name)
        return out[:n].copy()

    def step_sizes(self):
        out = np.zeros(3)
        self._f("get_step_sizes")(self.h, _p(out))
        return out


class UnOCPSolver(_SolverBase):
    """src/unocp/unocp_solver.cpp restated."""
    _prefix = "oracle_unocp_"

    def clear_line_search_filter(self):
        self.L.oracle_unocp_clear_line_search_filter(self.h)

    def is_feasible(self):
        return bool(self.L.oracle_unocp_is_feasible(self.h))

    def set_stage_threads(self, n):
        self.L.oracle_unocp_set_stage_threads(self.h, int(n))

    def get_constraint_data(self, name):
        out = np.zeros((self.N, 2 if name.startswith("acc_") else NC, NV))
        if self.L.oracle_unocp_get_constraint_data(self.h, name.encode(), _p(out)) < 0:
            raise ValueError(name)
        return out

    def get_unkkt(self, stage):
        Q = np.zeros((21, 21))
        res = np.zeros(35)
        self.L.oracle_unocp_get_unkkt(self.h, stage, _p(Q), _p(res))
        return Q.T.copy(), res

    def get_riccati(self, stage):
        Pqq, Pqv, Pvv = np.zeros((NV, NV)), np.zeros((NV, NV)), np.zeros((NV, NV))
        sq, sv, K, k = np.zeros(NV), np.zeros(NV), np.zeros((2 * NV, NV)), np.zeros(NV)
        self.L.oracle_unocp_get_riccati(self.h, stage, _p(Pqq), _p(Pqv), _p(Pvv), _p(sq), _p(sv), _p(K), _p(k))
        return dict(Pqq=Pqq.T.copy(), Pqv=Pqv.T.copy(), Pvv=Pvv.T.copy(), sq=sq, sv=sv, K=K.T.copy(), k=k)


class UnParNMPCSolver(_SolverBase):
    """src/unocp/unparnmpc_solver.cpp restated."""
    _prefix = "oracle_unparnmpc_"

    def init_backward_correction(self, t):
        self.L.oracle_unparnmpc_init_backward_correction(self.h, C.c_double(t))

    def clear_line_search_filter(self):
        self.L.oracle_unparnmpc_clear_line_search_filter(self.h)

    def get_kkt_inverse(self, stage):
        """35x35 KKT inverse of the last coarse update, order [lmd,gmm | a,q,v], numpy [row, col]."""
        out = np.zeros((35, 35))
        self.L.oracle_unparnmpc_get_kkt_inverse(self.h, stage, _p(out))
        return out.T.copy()

    def chol_info(self):
        return int(self.L.oracle_unparnmpc_chol_info(self.h))


def invert_unkkt(dt, Q):
    """SplitUnKKTMatrixInverter::invert on a 21x21 Q (numpy [row, col]); returns (info, 35x35 inverse)."""
    Qc = np.asfortranarray(np.asarray(Q, dtype=np.float64))
    out = np.zeros((35, 35))
    info = lib().oracle_invert_unkkt(C.c_double(dt), Qc.ctypes.data_as(_dp), _p(out))
    return info, out.T.copy()


class Batch:
    """A batch of independent oracle solvers driven with OpenMP over instances (BASELINE.md mode B)."""

    def __init__(self, problem, batch, kind="unocp"):
        self.kind = kind
        cls = UnOCPSolver if kind == "unocp" else UnParNMPCSolver
        self.solvers = [cls(problem) for _ in range(batch)]
        self.arr = (C.c_void_p * batch)(*[s.h for s in self.solvers])
        self.batch = batch
        self.L = lib()

    def update_solution(self, t, q0, v0, line_search=False, nthreads=1):
        q0, v0 = _vec(q0), _vec(v0)
        f = getattr(self.L, "oracle_%s_batch_update_solution" % self.kind)
        f(self.arr, self.batch, C.c_double(t), _p(q0), _p(v0), int(line_search), int(nthreads))

    def kkt_error(self, t, q0, v0, nthreads=1):
        q0, v0 = _vec(q0), _vec(v0)
        out = np.zeros(self.batch)
        f = getattr(self.L, "oracle_%s_batch_kkt" % self.kind)
        f(self.arr, self.batch, C.c_double(t), _p(q0), _p(v0), _p(out), int(nthreads))
        return out

    def _get(self, which, name):
        assert self.kind == "unocp"
        N = self.solvers[0].N
        out = np.zeros((self.batch, N + 1, NV))
        n = self.L.oracle_unocp_batch_get(self.arr, self.batch, int(which), name.encode(), _p(out))
        if n < 0:
            raise ValueError(name)
        return out[:, :n].copy()

    def get_solution(self, name):
        """(batch, stages, 7) of one solution field of every instance."""
        return self._get(0, name)

    def get_direction(self, name):
        return self._get(1, name)

    def step_sizes(self):
        out = np.zeros((self.batch, 3))
        self.L.oracle_unocp_batch_step_sizes(self.arr, self.batch, _p(out))
        return out


def task_space_3d_problem(N=30, T=1.5):
    """TaskSpace3DCost / TimeVaryingTaskSpace3DCost (src/cost/task_space_3d_cost.cpp) on the task_space_ocp robot set-up:
    position error of the end-effector frame, weights 1000 (task_enabled = 2)."""
    p = task_space_problem(N, T)
    p.task_enabled = 2
    for k in range(3, 6):
        p.task_q_weight[k] = 0.0
        p.task_qf_weight[k] = 0.0
    for i in range(7):          # a 3D position cost has rank 3 in q: a small posture weight keeps the stage Hessian definite
        p.q_weight[i] = 0.1     # (UnParNMPC factorises the full 21 x 21 stage Hessian)
        p.qf_weight[i] = 0.1
    return p


def task_evaluate_kind(q, ref12, kind):
    q, ref12 = _vec(q), _vec(ref12)
    diff, JJ = np.zeros(6), np.zeros((7, 6))
    lib().oracle_task_evaluate_kind(_p(q), _p(ref12), int(kind), _p(diff), _p(JJ))
    return diff, JJ.T.copy()


def task_evaluate(q, ref12):
    """diff_6d = log6(SE3_ref^-1 oMf) [lin; ang] and JJ = Jlog6 * J_frame(LOCAL) as numpy (6, 7)."""
    q, ref12 = _vec(q), _vec(ref12)
    diff, JJ = np.zeros(6), np.zeros((7, 6))
    lib().oracle_task_evaluate(_p(q), _p(ref12), _p(diff), _p(JJ))
    return diff, JJ.T.copy()


def frame_kinematics(q):
    """End-effector (frame 22) placement (R (3,3), p (3,)) and LOCAL frame Jacobian (6, 7)."""
    q = _vec(q)
    M, J = np.zeros(12), np.zeros((7, 6))
    lib().oracle_frame_kinematics(_p(q), _p(M), _p(J))
    return M[:9].reshape(3, 3).copy(), M[9:].copy(), J.T.copy()


def canon_acos(x):
    f = lib().oracle_canon_acos
    f.restype = C.c_double
    f.argtypes = [C.c_double]
    return f(float(x))

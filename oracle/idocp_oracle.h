/*
 * idocp_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 *
 * Plain-C restatement of idocp's Newton-step hot path for the fixed-base iiwa14
 * (UnOCPSolver / UnParNMPCSolver / UnLineSearch).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (idocp_b200/) never links, imports or executes it.
 *
 * PARITY UNPINNED at the pinocchio boundary: the reference (mayataka/idocp) cannot be
 * compiled here (Eigen, Boost, pinocchio, urdfdom are absent) and its
 * test-suite holds no golden vectors, so this oracle is validated by (i) finite
 * differences / an independent body-frame RNEA (oracle/np_mirror.py), (ii) the algebraic
 * identities the reference's own unit tests assert (tests/test_oracle_*.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Matrices are column-major (Eigen default): A[col*NV + row].
 */
#ifndef IDOCP_ORACLE_H_
#define IDOCP_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_NV 7          /* iiwa14: nq = nv = nu = 7 */
#define ORACLE_NC 6          /* constraint components of JointConstraintsFactory */

/* Problem description = what examples/iiwa14/{unocp_benchmark,config_space_ocp}.cpp build:
 * Robot limits + ConfigurationSpaceCost + JointConstraintsFactory constraints. */
typedef struct {
  int N;                      /* horizon stages */
  double T;                   /* horizon length */
  double q_ref[ORACLE_NV], v_ref[ORACLE_NV], u_ref[ORACLE_NV];
  double q_weight[ORACLE_NV], v_weight[ORACLE_NV], a_weight[ORACLE_NV], u_weight[ORACLE_NV];
  double qf_weight[ORACLE_NV], vf_weight[ORACLE_NV];
  double q_min[ORACLE_NV], q_max[ORACLE_NV], v_max[ORACLE_NV], u_max[ORACLE_NV];
  double barrier;             /* 1e-4 (joint_position_lower_limit.hpp:19-20) */
  double fraction_rate;       /* 0.995 */
  /* TimeVaryingTaskSpace6DCost (src/cost/time_varying_task_space_6d_cost.cpp) on the end
   * effector frame, reference = circle of examples/iiwa14/task_space_ocp.cpp:21-46.
   * Disabled when task_enabled == 0; task_enabled == 2 selects TaskSpace3DCost / TimeVaryingTaskSpace3DCost
   * (src/cost/task_space_3d_cost.cpp): position error only, task_q_weight[0..2] = q_3d_weight. */
  int task_enabled;
  double task_q_weight[6], task_qf_weight[6];  /* [position xyz, rotation xyz] = the arguments of set_q_6d_weight /
                                                  set_qf_6d_weight (time_varying_task_space_6d_cost.cpp:43-58) */
  double task_center[3], task_radius, task_t0, task_tf; /* unused by the engine: the reference is a host-sampled table */
  double task_rot_ref[9];
  /* JointAccelerationLowerLimit / JointAccelerationUpperLimit (src/constraints/joint_acceleration_*_limit.cpp) with their
   * constructor arguments amin / amax; [0] lower, [1] upper */
  int enable_acc[2];
  double a_min[ORACLE_NV], a_max[ORACLE_NV];
} oracle_problem_t;

void oracle_problem_default(oracle_problem_t* p);

/* robot/  (pinocchio call sites: include/idocp/robot/robot.hxx:444-500) */
void oracle_rnea(const double* q, const double* v, const double* a, double* tau);
void oracle_rnea_derivatives(const double* q, const double* v, const double* a,
                             double* dtau_dq, double* dtau_dv, double* dtau_da);

/* UnOCPSolver (src/unocp/unocp_solver.cpp) -- one instance */
typedef struct oracle_unocp oracle_unocp_t;
oracle_unocp_t* oracle_unocp_create(const oracle_problem_t* p);
void oracle_unocp_destroy(oracle_unocp_t* o);
int  oracle_unocp_set_solution(oracle_unocp_t* o, const char* name, const double* value);
void oracle_unocp_init_constraints(oracle_unocp_t* o);
void oracle_unocp_update_solution(oracle_unocp_t* o, double t, const double* q, const double* v,
                                  int line_search);
void oracle_unocp_compute_kkt_residual(oracle_unocp_t* o, double t, const double* q, const double* v);
double oracle_unocp_kkt_error(oracle_unocp_t* o);
void oracle_unocp_clear_line_search_filter(oracle_unocp_t* o);
int  oracle_unocp_is_feasible(oracle_unocp_t* o);
/* field in {"lmd","gmm","q","v","a","u","beta"}: out[(N+1)*NV] (a,u,beta: N*NV) */
int  oracle_unocp_get_solution(const oracle_unocp_t* o, const char* name, double* out);
/* field in {"dlmd","dgmm","dq","dv","da","du","dbeta"} */
int  oracle_unocp_get_direction(const oracle_unocp_t* o, const char* name, double* out);
/* field in {"slack","dual","residual","duality","dslack","ddual"}: out[N*NC*NV], inactive rows = 0;
 * "acc_" + field: the two acceleration-limit components, out[N*2*NV] */
int  oracle_unocp_get_constraint_data(const oracle_unocp_t* o, const char* name, double* out);
/* last step sizes: out[0]=primal (after line search), out[1]=dual, out[2]=max primal (before) */
void oracle_unocp_get_step_sizes(const oracle_unocp_t* o, double* out);
/* condensed stage data of the last linearisation: Q[21*21] col-major (order a,q,v), res[35] = [Fq,Fv,la,lq,lv] */
void oracle_unocp_get_unkkt(const oracle_unocp_t* o, int stage, double* Q, double* res);
/* Riccati factorisation of a stage: Pqq,Pqv,Pvv [49 each], sq,sv [7 each], K [7x14], k[7] (K,k: stage<N) */
void oracle_unocp_get_riccati(const oracle_unocp_t* o, int stage, double* Pqq, double* Pqv, double* Pvv,
                              double* sq, double* sv, double* K, double* k);

/* batch drivers (OpenMP over instances = BASELINE.md "mode B") */
void oracle_unocp_batch_update_solution(oracle_unocp_t** os, int batch, double t, const double* q0,
                                        const double* v0, int line_search, int nthreads);
void oracle_unocp_batch_kkt(oracle_unocp_t** os, int batch, double t, const double* q0,
                            const double* v0, double* kkt_out, int nthreads);
int  oracle_unocp_batch_get(oracle_unocp_t** os, int batch, int which, const char* name, double* out);
void oracle_unocp_batch_step_sizes(oracle_unocp_t** os, int batch, double* out);
/* reference threading (OpenMP over stages inside one instance = "mode A", unocp_benchmark.cpp:42) */
void oracle_unocp_set_stage_threads(oracle_unocp_t* o, int nthreads);

/* UnParNMPCSolver (src/unocp/unparnmpc_solver.cpp) */
typedef struct oracle_unparnmpc oracle_unparnmpc_t;
oracle_unparnmpc_t* oracle_unparnmpc_create(const oracle_problem_t* p);
void oracle_unparnmpc_destroy(oracle_unparnmpc_t* o);
int  oracle_unparnmpc_set_solution(oracle_unparnmpc_t* o, const char* name, const double* value);
void oracle_unparnmpc_init_constraints(oracle_unparnmpc_t* o);
void oracle_unparnmpc_init_backward_correction(oracle_unparnmpc_t* o, double t);
void oracle_unparnmpc_update_solution(oracle_unparnmpc_t* o, double t, const double* q, const double* v,
                                      int line_search);
void oracle_unparnmpc_compute_kkt_residual(oracle_unparnmpc_t* o, double t, const double* q, const double* v);
double oracle_unparnmpc_kkt_error(oracle_unparnmpc_t* o);
void oracle_unparnmpc_clear_line_search_filter(oracle_unparnmpc_t* o);
int  oracle_unparnmpc_get_solution(const oracle_unparnmpc_t* o, const char* name, double* out);
int  oracle_unparnmpc_get_direction(const oracle_unparnmpc_t* o, const char* name, double* out);
void oracle_unparnmpc_get_step_sizes(const oracle_unparnmpc_t* o, double* out);
/* 35x35 inverse of [[0 F],[F^T Q]] of a stage (column-major, order [lmd,gmm | a,q,v]) */
void oracle_unparnmpc_get_kkt_inverse(const oracle_unparnmpc_t* o, int stage, double* out);
int  oracle_unparnmpc_chol_info(const oracle_unparnmpc_t* o);
/* SplitUnKKTMatrixInverter::invert (unocp/split_unkkt_matrix_inverter.hxx:40-79) on a 21x21 Q */
int  oracle_invert_unkkt(double dt, const double* Q, double* Kinv);
void oracle_unparnmpc_batch_update_solution(oracle_unparnmpc_t** os, int batch, double t, const double* q0,
                                            const double* v0, int line_search, int nthreads);
void oracle_unparnmpc_batch_kkt(oracle_unparnmpc_t** os, int batch, double t, const double* q0,
                                const double* v0, double* kkt_out, int nthreads);

/* TimeVaryingTaskSpace6DCost: host-sampled reference table [(N+1)][12] (R row-major, p), see idocp_oracle.c */
void oracle_unocp_set_task_ref(oracle_unocp_t* o, const double* table);
void oracle_unparnmpc_set_task_ref(oracle_unparnmpc_t* o, const double* table);
void oracle_task_evaluate(const double* q, const double* ref12, double* diff6, double* JJ);
/* kind = problem.task_enabled: 1 the 6D cost, 2 TaskSpace3DCost (diff_3d / J_3d in the first three entries / rows) */
void oracle_task_evaluate_kind(const double* q, const double* ref12, int kind, double* diff6, double* JJ);
void oracle_frame_kinematics(const double* q, double* oMf12, double* J);
double oracle_canon_acos(double x);

/* splitmix64 counter-based generator shared by oracle, bench and tests */
double oracle_splitmix_uniform(unsigned long long seed, unsigned long long index);

#ifdef __cplusplus
}
#endif
#endif

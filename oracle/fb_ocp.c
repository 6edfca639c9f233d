/* oracle/fb_ocp.c -- TEST INFRASTRUCTURE: CPU oracle of idocp's OCPSolver (floating base, contacts, impulses,
 * switching constraints) for ANYmal, SURVEY.md §8 row a12.  Plain-C restatement, file by file, of
 *
 *   OCPSolver::updateSolution / computeKKTResidual / KKTError / initConstraints / setSolution
 *                                                          src/ocp/ocp_solver.cpp:60-316
 *   OCPLinearizer (stage dispatch, q_prev, integrateSolution) include/idocp/ocp/ocp_linearizer.hxx:113-248,
 *                                                          src/ocp/ocp_linearizer.cpp:40-221
 *   SplitOCP / TerminalOCP / ImpulseSplitOCP               ocp/split_ocp.hxx, terminal_ocp.hxx, impulse/impulse_split_ocp.hxx
 *   stateequation (forward Euler, SE(3) blocks)            ocp/state_equation.hxx:11-110, impulse/impulse_state_equation.hxx
 *   ContactDynamics / ImpulseDynamicsForwardEuler          ocp/contact_dynamics.hxx:48-231, impulse/impulse_dynamics_forward_euler.hxx
 *   ForwardSwitchingConstraint                             ocp/forward_switching_constraint.hxx:27-81
 *   SplitRiccatiFactorizer (+constrained), BackwardRiccatiRecursionFactorizer, Impulse* twins
 *                                                          ocp/split_riccati_factorizer.hxx, backward_riccati_recursion_factorizer.hxx
 *   RiccatiRecursionSolver                                 src/ocp/riccati_recursion_solver.cpp:48-251
 *   cost: configuration-space cost with a time-varying reference (TrottingConfigurationSpaceCost etc.:
 *   src/cost/trotting_configuration_space_cost.cpp:241-375; the reference q_ref(t) is a host-side function of the
 *   stage time and enters as a per-stage table), ContactForceCost (src/cost/contact_force_cost.cpp:121-228)
 *   constraints: Joint{Position,Velocity,Torques}{Lower,Upper}Limit, LinearizedFrictionCone,
 *   LinearizedImpulseFrictionCone (src/constraints/ *.cpp), pdipm.hxx
 *
 * "Parity unpinned" (no buildable reference, no golden vectors): validated by the identities of the reference's
 * unit tests and by convergence of examples/anymal/anymal_trotting.cpp (tests/test_oracle_fb_ocp.py).
 * Dense algebra: every output element is an ascending-index fma chain (mm()), so any parallel mapping of the
 * elements reproduces the bits. */
#include <stdio.h>
#include <stdlib.h>

#include "fb_robot.h"
#include "hybrid_oracle.h"

#define NV FB_NV
#define NQ FB_NQ
#define NU FB_NU
#define NX 36
#define NVF 30
#define MAXF FB_MAXF
#define NPASS 6

enum { K_GRID = 0, K_IMPULSE = 1, K_AUX = 2, K_LIFT = 3, K_TERMINAL = 4 };
enum { C_POS_LO = 0, C_POS_UP, C_VEL_LO, C_VEL_UP, C_TRQ_LO, C_TRQ_UP, C_FRICTION, C_IMPULSE_FRICTION, C_ACC_LO, C_ACC_UP, C_DISTANCE, NCOMP };

typedef struct {
  double T;
  int N, max_num_impulse;
  double q_weight[NV], v_weight[NV], a_weight[NV], qf_weight[NV], vf_weight[NV], qi_weight[NV], vi_weight[NV], dvi_weight[NV];
  double f_weight[MAXF], f_ref[MAXF], fi_weight[MAXF], fi_ref[MAXF];
  double q_min[NU], q_max[NU], v_max[NU], u_max[NU];
  double mu, barrier, fraction_rate;
  int enable[8];
  /* appended (idocp_b200_fb_problem keeps its first members): FrictionCone / ImpulseFrictionCone (the nonlinear cones,
   * src/constraints/friction_cone.cpp, impulse_friction_cone.cpp) instead of the linearised ones; JointAcceleration{Lower,
   * Upper}Limit (joint_acceleration_*_limit.cpp) with their amin / amax constructor arguments */
  int cone_nonlinear[2];
  int enable_acc[2];
  double a_min[NU], a_max[NU];
  /* ContactDistance (src/constraints/contact_distance.cpp): the height of the contact frame of every contact that is NOT
   * active stays positive; position level.  1 = the reference literally: the gradient / Hessian rows are row 2 of
   * getFrameJacobian, which is the LOCAL-frame Jacobian (robot.hxx:182-188), i.e. NOT the derivative of the world height
   * unless the foot frame is level -- the Newton iteration diverges on the trot (tests/test_oracle_fb_ocp.py);
   * 2 = the consistent variant, row 2 of R_frame J_linear = d z / d q */
  int enable_distance;
} oracle_fb_problem_t;
static inline int comp_enabled(const oracle_fb_problem_t* p, int c) {
  return c < C_ACC_LO ? p->enable[c] : (c == C_DISTANCE ? p->enable_distance : p->enable_acc[c - C_ACC_LO]);
}
static inline int is_cone(int c) { return c == C_FRICTION || c == C_IMPULSE_FRICTION; }
/* rows per contact: 5 (linearised) or 2 (normal force, cone) */
static inline int cone_rows(const oracle_fb_problem_t* p, int c) { return p->cone_nonlinear[c - C_FRICTION] ? 2 : 5; }

typedef struct { double slack[20], dual[20], residual[20], duality[20], dslack[20], ddual[20]; } cdata_t;
static inline int comp_dim(int c) { return (c == C_FRICTION || c == C_IMPULSE_FRICTION) ? 20 : (c == C_DISTANCE ? FB_NC : NU); }   /* storage */
#define comp_rows(p, c) (is_cone(c) ? FB_NC * cone_rows(p, c) : ((c) == C_DISTANCE ? FB_NC : NU))                                /* live rows */

typedef struct {
  /* --- SplitSolution / ImpulseSplitSolution (a = dv at an impulse) --- */
  double lmd[NV], gmm[NV], q[NQ], v[NV], a[NV], u[NU], beta[NV], nu_passive[NPASS], f[FB_NC][3], mu[FB_NC][3], xi[MAXF];
  int active[FB_NC], dimf, imp_active[FB_NC], dimi;
  double cpoints[FB_NC][3], ipoints[FB_NC][3];
  double ref_q[NQ], ref_v[NV];
  /* --- SplitDirection --- */
  double dlmd[NV], dgmm[NV], dq[NV], dv[NV], du[NU], daf[NVF], dbetamu[NVF], dnu_passive[NPASS], dxi[MAXF];
  /* --- ConstraintsData --- */
  int cstage, cactive[NCOMP];
  cdata_t c[NCOMP];
  double cdJ[FB_NC][NV], cdz[FB_NC];   /* ContactDistance: data.J[i].row(2) (LOCAL frame Jacobian) and the frame height */
  /* --- SplitKKTResidual --- */
  double lq[NV], lv[NV], la[NV], lf[MAXF], lu_passive[NPASS], lu[NU], Fq[NV], Fv[NV], Fq_prev[6], P[MAXF];
  /* --- SplitKKTMatrix (only the blocks the path touches) --- */
  double Qxx[NX * NX], Qxu[NX * NV], Quu[NV * NV], Qaa[NV], Qff[MAXF * MAXF];
  double Fqq6[36], Fqv6[36], Fqq_prev6[36], Fqq_inv[36], Fqq_prev_inv[36], Fvq[NV * NV], Fvv[NV * NV], Fvu[NV * NU];
  /* --- ContactDynamicsData / ImpulseDynamicsForwardEulerData --- */
  double IDC[NVF], dIDCdqv[NVF * NX], Mm[NV * NV], dCda[MAXF * NV], MJtJinv[NVF * NVF], MJ_dIDC[NVF * NX], MJ_IDC[NVF],
      Qafqv[NVF * NX], Qafu[NVF * NV], laf[NVF];
  /* --- SplitStateConstraintJacobian --- */
  double Pq[MAXF * NV], Phix[MAXF * NX], Phia[MAXF * NV], Phiu[MAXF * NU];
  /* --- LQRStateFeedbackPolicy, SplitRiccatiFactorization, SplitConstrainedRiccatiFactorization --- */
  double K[NU * NX], k[NU], Pqq[NV * NV], Pqv[NV * NV], Pvv[NV * NV], sq[NV], sv[NV], cM[MAXF * NX], cm[MAXF];
  double max_primal, max_dual, kkt_sq;
  int chol_info;
  /* --- LineSearch: trial point s + alpha d (line_search.hpp:130-158) and its stage cost / violation --- */
  double tq[NQ], tv[NV], ta[NV], tu[NU], tf[FB_NC][3], ls_cost, ls_viol;
} stage_t;

typedef struct {
  int kind, index, slot;
  double t, dt;
  int phase;      /* contact phase (GRID/AUX/LIFT/TERMINAL) or impulse index (IMPULSE) */
  int cstage;     /* createConstraintsData argument */
  int sw_impulse; /* >= 0: the switching constraint of this impulse is imposed here */
  double dt_next;
  int ls_next;    /* chain position whose (q, v) closes the state equation in LineSearch::computeCostAndViolation */
  int ls_impulse; /* >= 0: impulse whose switching residual the line search adds on this stage */
  double ls_dt_next;
} elem_t;

#define MAX_ELEMS (HY_MAX_N + 1 + 3 * HY_MAX_EVENTS)

typedef struct oracle_fb_ocp {
  oracle_fb_problem_t p;
  const oracle_contact_sequence_t* cs;
  int n_slots, n_elems;
  stage_t* slots;
  elem_t elems[MAX_ELEMS];
  oracle_discretization_t disc;
  double primal_step, dual_step;
  double q0_prev[NQ];
  int nthreads;
  int filter_n;
  double filter_cost[256], filter_viol[256];
} oracle_fb_ocp_t;

/* ---------------------------------------------------------------------------------------------- */
/* dense helper: C (m x n, ldc) {=, +=, -=} A (m x k) B (k x n), arbitrary strides, ascending-k fma chain */
/* ---------------------------------------------------------------------------------------------- */
enum { MM_SET = 0, MM_ADD = 1, MM_SUB = 2 };
static void mm(int mode, int m, int n, int k, const double* A, int ars, int acs, const double* B, int brs, int bcs,
               double* C, int ldc) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double acc;
      int l0 = 0;
      if (mode == MM_SET) {
        if (k == 0) { C[i * ldc + j] = 0.0; continue; }
        acc = A[i * ars] * B[j * bcs];
        l0 = 1;
      } else {
        acc = C[i * ldc + j];
      }
      if (mode == MM_SUB)
        for (int l = l0; l < k; ++l) acc = fma(-A[i * ars + l * acs], B[l * brs + j * bcs], acc);
      else
        for (int l = l0; l < k; ++l) acc = fma(A[i * ars + l * acs], B[l * brs + j * bcs], acc);
      C[i * ldc + j] = acc;
    }
}
/* y (m) op= A (m x k) x */
static void mv(int mode, int m, int k, const double* A, int ars, int acs, const double* x, double* y) {
  mm(mode, m, 1, k, A, ars, acs, x, 1, 1, y, 1);
}
static double sqnorm_n(const double* x, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc = fma(x[i], x[i], acc);
  return acc;
}
static double l1norm_n(const double* x, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc += fabs(x[i]);
  return acc;
}

/* ---------------------------------------------------------------------------------------------- */
/* constraints/ (pdipm.hxx, joint_*_limit.cpp, linearized_friction_cone.cpp, constraints_data.hpp)  */
/* ---------------------------------------------------------------------------------------------- */
static void set_constraint_stage(const oracle_fb_problem_t* p, stage_t* st, int time_stage) {
  st->cstage = time_stage;
  const int pos = time_stage >= 2, vel = time_stage >= 1, acc = time_stage >= 0, imp = time_stage <= -1;
  st->cactive[C_DISTANCE] = pos && p->enable_distance;   /* KinematicsLevel::PositionLevel (contact_distance.cpp:31-33) */
  st->cactive[C_ACC_LO] = acc && p->enable_acc[0];
  st->cactive[C_ACC_UP] = acc && p->enable_acc[1];
  st->cactive[C_POS_LO] = pos && p->enable[C_POS_LO];
  st->cactive[C_POS_UP] = pos && p->enable[C_POS_UP];
  st->cactive[C_VEL_LO] = vel && p->enable[C_VEL_LO];
  st->cactive[C_VEL_UP] = vel && p->enable[C_VEL_UP];
  st->cactive[C_TRQ_LO] = acc && p->enable[C_TRQ_LO];
  st->cactive[C_TRQ_UP] = acc && p->enable[C_TRQ_UP];
  st->cactive[C_FRICTION] = acc && p->enable[C_FRICTION];
  st->cactive[C_IMPULSE_FRICTION] = imp && p->enable[C_IMPULSE_FRICTION];
}
/* frictionConeResidual (linearized_friction_cone.hpp:72-85); nonlinear: normalForceResidual, frictionConeResidual
 * (friction_cone.hpp:72-82) */
static inline void friction_residual(double mu, int nonlinear, const double* f, double* r) {
  if (nonlinear) {
    r[0] = -f[2];
    r[1] = fma(-((mu * mu) * f[2]), f[2], fma(f[1], f[1], f[0] * f[0]));
    r[2] = r[3] = r[4] = 0.0;
    return;
  }
  const double s = mu * f[2] / 1.41421356237309514547e+00;
  r[0] = -f[2];
  r[1] = f[0] - s;
  r[2] = -f[0] - s;
  r[3] = f[1] - s;
  r[4] = -f[1] - s;
}
/* Jac_ of the cone (linearized_friction_cone.cpp:25-29), row e; nonlinear: row 0 = d(-fz)/df, row 1 = data.r[i] =
 * (2 fx, 2 fy, -2 mu^2 fz) (friction_cone.cpp:109-111) */
static inline void friction_jac_row(double mu, int nonlinear, const double* f, int e, double* row) {
  if (nonlinear) {
    if (e == 0) { row[0] = 0.0; row[1] = 0.0; row[2] = -1.0; }
    else { row[0] = 2.0 * f[0]; row[1] = 2.0 * f[1]; row[2] = -(((2.0 * mu) * mu) * f[2]); }
    return;
  }
  const double m = -(mu / 1.41421356237309514547e+00);
  const double J[5][3] = {{0, 0, -1}, {1, 0, m}, {-1, 0, m}, {0, 1, m}, {0, -1, m}};
  row[0] = J[e][0]; row[1] = J[e][1]; row[2] = J[e][2];
}
/* margin g of joint-limit component c at joint j: slack is initialised to g, residual = -g + slack */
static inline double limit_margin(const oracle_fb_problem_t* p, int c, const stage_t* s, int j) {
  switch (c) {
    case C_POS_LO: return s->q[7 + j] - p->q_min[j];
    case C_POS_UP: return p->q_max[j] - s->q[7 + j];
    case C_VEL_LO: return s->v[6 + j] - (-p->v_max[j]);
    case C_VEL_UP: return p->v_max[j] - s->v[6 + j];
    case C_TRQ_LO: return s->u[j] - (-p->u_max[j]);
    case C_ACC_LO: return s->a[6 + j] - p->a_min[j];
    case C_ACC_UP: return p->a_max[j] - s->a[6 + j];
    default:       return p->u_max[j] - s->u[j];
  }
}
static inline double limit_residual_at(const oracle_fb_problem_t* p, int c, const double* q, const double* v, const double* a,
                                       const double* u, int j, double slack) {
  switch (c) {
    case C_POS_LO: return p->q_min[j] - q[7 + j] + slack;
    case C_POS_UP: return q[7 + j] - p->q_max[j] + slack;
    case C_VEL_LO: return (-p->v_max[j]) - v[6 + j] + slack;
    case C_VEL_UP: return v[6 + j] - p->v_max[j] + slack;
    case C_TRQ_LO: return (-p->u_max[j]) - u[j] + slack;
    case C_ACC_LO: return p->a_min[j] - a[6 + j] + slack;
    case C_ACC_UP: return a[6 + j] - p->a_max[j] + slack;
    default:       return u[j] - p->u_max[j] + slack;
  }
}
static inline double limit_residual(const oracle_fb_problem_t* p, int c, const stage_t* s, int j, double slack) {
  return limit_residual_at(p, c, s->q, s->v, s->a, s->u, j, slack);
}
static inline double limit_sign(int c) { return (c & 1) ? 1.0 : -1.0; } /* lower: -, upper: + */

/* Constraints::setSlackAndDual (constraints.hxx) + pdipm::SetSlackAndDualPositive (pdipm.hxx:13-23).
 * The friction cones initialise every contact, active or not (linearized_friction_cone.cpp:102-111). */
static void set_slack_and_dual(const oracle_fb_problem_t* p, stage_t* st) {
  for (int c = 0; c < NCOMP; ++c) {
    cdata_t* d = &st->c[c];
    memset(d, 0, sizeof(*d));
    if (!st->cactive[c]) continue;
    const int n = comp_rows(p, c);
    fb_kin_t kin;
    if (c == C_DISTANCE) fb_forward_kinematics(st->q, NULL, NULL, &kin);   /* robot.updateFrameKinematics(s.q) */
    for (int j = 0; j < n; ++j) {
      double sl;
      if (c == C_DISTANCE) {   /* every contact, active or not (contact_distance.cpp:62-70) */
        double P[3];
        fb_contact_point(&kin, j, P);
        sl = P[2];
      } else if (is_cone(c)) {
        const int rpc = cone_rows(p, c);
        double r[5];
        friction_residual(p->mu, rpc == 2, st->f[j / rpc], r);
        sl = -r[j % rpc];
      } else {
        sl = limit_margin(p, c, st, j);
      }
      for (int guard = 0; sl < p->barrier && guard < (1 << 20); ++guard) sl += p->barrier;   /* bound: see k_fb_init_constraints */
      d->slack[j] = sl;
      d->dual[j] = p->barrier / sl;
    }
  }
}
/* computePrimalAndDualResidual of every live component */
static void primal_dual_residual(const oracle_fb_problem_t* p, stage_t* st) {
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c] || c == C_DISTANCE) continue;   /* ContactDistance needs the kinematics: contact_distance_stage() */
    cdata_t* d = &st->c[c];
    if (is_cone(c)) {
      const int rpc = cone_rows(p, c);
      for (int i = 0; i < FB_NC; ++i) {
        double r[5];
        if (st->active[i]) friction_residual(p->mu, rpc == 2, st->f[i], r);
        for (int e = 0; e < rpc; ++e) {
          const int j = rpc * i + e;
          if (st->active[i]) {
            d->residual[j] = r[e] + d->slack[j];
            d->duality[j] = d->slack[j] * d->dual[j] - p->barrier;
          } else {
            d->residual[j] = 0.0;
            d->duality[j] = 0.0;
          }
        }
      }
    } else {
      for (int j = 0; j < NU; ++j) {
        d->residual[j] = limit_residual(p, c, st, j, d->slack[j]);
        d->duality[j] = d->slack[j] * d->dual[j] - p->barrier;
      }
    }
  }
}
static inline double* limit_grad(stage_t* st, int c) {
  return c <= C_POS_UP ? st->lq + 6 : (c <= C_VEL_UP ? st->lv + 6 : (c >= C_ACC_LO ? st->la + 6 : st->lu));
}

/* Constraints::augmentDualResidual; dt = 1 at an impulse (no dt argument there) */
static void augment_dual_residual(const oracle_fb_problem_t* p, stage_t* st, double dt) {
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c] || c == C_DISTANCE) continue;
    const cdata_t* d = &st->c[c];
    if (is_cone(c)) {
      const int rpc = cone_rows(p, c);
      int k = 0;
      for (int i = 0; i < FB_NC; ++i) {
        if (!st->active[i]) continue;
        for (int x = 0; x < 3; ++x) {
          double acc = 0.0;
          for (int e = 0; e < rpc; ++e) {
            double row[3];
            friction_jac_row(p->mu, rpc == 2, st->f[i], e, row);
            acc = (e == 0) ? row[x] * d->dual[rpc * i + e] : fma(row[x], d->dual[rpc * i + e], acc);
          }
          st->lf[3 * k + x] += dt * acc;
        }
        ++k;
      }
    } else {
      double* l = limit_grad(st, c);
      const double sg = limit_sign(c);
      for (int j = 0; j < NU; ++j) l[j] += sg * (dt * d->dual[j]);
    }
  }
}
/* Constraints::condenseSlackAndDual */
static void condense_slack_and_dual(const oracle_fb_problem_t* p, stage_t* st, double dt) {
  primal_dual_residual(p, st);
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c] || c == C_DISTANCE) continue;
    cdata_t* d = &st->c[c];
    if (is_cone(c)) {
      const int rpc = cone_rows(p, c);
      int k = 0;
      for (int i = 0; i < FB_NC; ++i) {
        if (!st->active[i]) continue;
        double r5[5], w5[5], Jr[5][3];
        for (int e = 0; e < rpc; ++e) {
          const int j = rpc * i + e;
          const double rs = 1.0 / d->slack[j];
          r5[e] = fma(d->dual[j], d->residual[j], -d->duality[j]) * rs;
          w5[e] = d->dual[j] * rs;
          friction_jac_row(p->mu, rpc == 2, st->f[i], e, Jr[e]);
        }
        for (int x = 0; x < 3; ++x) {
          double acc = Jr[0][x] * r5[0];
          for (int e = 1; e < rpc; ++e) acc = fma(Jr[e][x], r5[e], acc);
          st->lf[3 * k + x] += dt * acc;
          for (int y = 0; y < 3; ++y) {
            double h = Jr[0][x] * (w5[0] * Jr[0][y]);
            for (int e = 1; e < rpc; ++e) h = fma(Jr[e][x], w5[e] * Jr[e][y], h);
            st->Qff[(3 * k + x) * MAXF + 3 * k + y] += dt * h;
          }
        }
        ++k;
      }
    } else {
      double* l = limit_grad(st, c);
      const double sg = limit_sign(c);
      for (int j = 0; j < NU; ++j) {
        const double rs = 1.0 / d->slack[j];
        const double h = (dt * d->dual[j]) * rs;
        if (c <= C_POS_UP) st->Qxx[(6 + j) * NX + 6 + j] += h;
        else if (c <= C_VEL_UP) st->Qxx[(NV + 6 + j) * NX + NV + 6 + j] += h;
        else if (c >= C_ACC_LO) st->Qaa[6 + j] += h;
        else st->Quu[(6 + j) * NV + 6 + j] += h;
        l[j] += sg * ((dt * fma(d->dual[j], d->residual[j], -d->duality[j])) * rs);
      }
    }
  }
}
/* Constraints::computeSlackAndDualDirection */
static void slack_dual_direction(const oracle_fb_problem_t* p, stage_t* st) {
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c]) continue;
    cdata_t* d = &st->c[c];
    if (c == C_DISTANCE) {   /* contact_distance.cpp:112-131: the rows of active contacts drift like the inactive cone rows */
      for (int i = 0; i < FB_NC; ++i) {
        d->dslack[i] = 1.0; d->ddual[i] = 1.0;
        if (st->active[i]) continue;
        double acc = st->cdJ[i][0] * st->dq[0];
        for (int l = 1; l < NV; ++l) acc = fma(st->cdJ[i][l], st->dq[l], acc);
        d->dslack[i] = acc - d->residual[i];
        d->ddual[i] = -fma(d->dual[i], d->dslack[i], d->duality[i]) / d->slack[i];
      }
    } else if (is_cone(c)) {
      const int rpc = cone_rows(p, c);
      for (int j = 0; j < FB_NC * rpc; ++j) { d->dslack[j] = 1.0; d->ddual[j] = 1.0; }
      int k = 0;
      for (int i = 0; i < FB_NC; ++i) {
        if (!st->active[i]) continue;
        const double* df = st->daf + NV + 3 * k;
        for (int e = 0; e < rpc; ++e) {
          const int j = rpc * i + e;
          double row[3];
          friction_jac_row(p->mu, rpc == 2, st->f[i], e, row);
          const double Jdf = fma(row[2], df[2], fma(row[1], df[1], row[0] * df[0]));
          d->dslack[j] = -Jdf - d->residual[j];
          d->ddual[j] = -fma(d->dual[j], d->dslack[j], d->duality[j]) / d->slack[j];
        }
        ++k;
      }
    } else {
      const double* dx = c <= C_POS_UP ? st->dq + 6 : (c <= C_VEL_UP ? st->dv + 6 : (c >= C_ACC_LO ? st->daf + 6 : st->du));
      for (int j = 0; j < NU; ++j) {
        d->dslack[j] = ((c & 1) ? -dx[j] : dx[j]) - d->residual[j];
        d->ddual[j] = -fma(d->dual[j], d->dslack[j], d->duality[j]) / d->slack[j];
      }
    }
  }
}
static double fraction_to_boundary(double rate, int n, const double* vec, const double* dvec) {
  double mn = 1.0;
  for (int i = 0; i < n; ++i) {
    const double f = -rate * (vec[i] / dvec[i]);
    if (f > 0 && f < 1) {
      if (f < mn) mn = f;
    }
  }
  return mn;
}
static void max_step_sizes(const oracle_fb_problem_t* p, stage_t* st) {
  double mp = 1.0, md = 1.0;
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c]) continue;
    const double a = fraction_to_boundary(p->fraction_rate, comp_rows(p, c), st->c[c].slack, st->c[c].dslack);
    const double b = fraction_to_boundary(p->fraction_rate, comp_rows(p, c), st->c[c].dual, st->c[c].ddual);
    if (a < mp) mp = a;
    if (b < md) md = b;
  }
  st->max_primal = mp;
  st->max_dual = md;
}
static double constraints_sqnorm(const stage_t* st) {
  double e = 0.0;
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c]) continue;
    e += sqnorm_n(st->c[c].residual, comp_dim(c)) + sqnorm_n(st->c[c].duality, comp_dim(c));
  }
  return e;
}

/* ---------------------------------------------------------------------------------------------- */
/* cost/                                                                                           */
/* ---------------------------------------------------------------------------------------------- */
/* configuration-space cost with reference (ref_q, ref_v): gradient (and Hessian) wrt q through
 * J_qdiff = dSubtractdConfigurationPlus(q, q_ref) (6x6 base block, identity on the joints). */
/* mode 0: gradient (compute*CostDerivatives), mode 2: Hessian (compute*CostHessian) */
static void cost_derivatives(const oracle_fb_problem_t* p, stage_t* st, int kind, double dt, int mode) {
  const double* wq = kind == K_TERMINAL ? p->qf_weight : (kind == K_IMPULSE ? p->qi_weight : p->q_weight);
  const double* wv = kind == K_TERMINAL ? p->vf_weight : (kind == K_IMPULSE ? p->vi_weight : p->v_weight);
  const double* wa = kind == K_IMPULSE ? p->dvi_weight : p->a_weight;
  const double sc = (kind == K_TERMINAL || kind == K_IMPULSE) ? 1.0 : dt;
  double qdiff[NV], J6[36], g[6];
  fb_subtract(st->q, st->ref_q, qdiff);
  fb_dsubtract_dplus(st->q, st->ref_q, J6);
  if (mode == 0) {
  for (int i = 0; i < 6; ++i) g[i] = wq[i] * qdiff[i];
  for (int r = 0; r < 6; ++r) {
    double acc = J6[r] * g[0];
    for (int k = 1; k < 6; ++k) acc = fma(J6[6 * k + r], g[k], acc);
    st->lq[r] += sc * acc;
  }
  for (int j = 6; j < NV; ++j) st->lq[j] += sc * (wq[j] * qdiff[j]);
  for (int j = 0; j < NV; ++j) st->lv[j] += sc * (wv[j] * (st->v[j] - st->ref_v[j]));
  if (kind != K_TERMINAL)
    for (int j = 0; j < NV; ++j) st->la[j] += sc * (wa[j] * st->a[j]);
  /* ContactForceCost */
  if (kind != K_TERMINAL) {
    const double* fw = kind == K_IMPULSE ? p->fi_weight : p->f_weight;
    const double* fr = kind == K_IMPULSE ? p->fi_ref : p->f_ref;
    int k = 0;
    for (int i = 0; i < FB_NC; ++i) {
      if (!st->active[i]) continue;
      for (int x = 0; x < 3; ++x) st->lf[3 * k + x] += sc * (fw[3 * i + x] * (st->f[i][x] - fr[3 * i + x]));
      ++k;
    }
  }
  return;
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      double acc = J6[r] * (wq[0] * J6[c]);
      for (int k = 1; k < 6; ++k) acc = fma(J6[6 * k + r], wq[k] * J6[6 * k + c], acc);
      st->Qxx[r * NX + c] += sc * acc;
    }
  for (int j = 6; j < NV; ++j) st->Qxx[j * NX + j] += sc * wq[j];
  for (int j = 0; j < NV; ++j) st->Qxx[(NV + j) * NX + NV + j] += sc * wv[j];
  if (kind != K_TERMINAL) {
    for (int j = 0; j < NV; ++j) st->Qaa[j] += sc * wa[j];
    const double* fw = kind == K_IMPULSE ? p->fi_weight : p->f_weight;
    int k = 0;
    for (int i = 0; i < FB_NC; ++i) {
      if (!st->active[i]) continue;
      for (int x = 0; x < 3; ++x) st->Qff[(3 * k + x) * MAXF + 3 * k + x] += sc * fw[3 * i + x];
      ++k;
    }
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* state equation (ocp/state_equation.hxx:11-63, impulse/impulse_state_equation.hxx:11-60)          */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { const double *lmd, *gmm, *q, *v; } next_t;

static void zero_residual(stage_t* st) {
  memset(st->lq, 0, sizeof(st->lq)); memset(st->lv, 0, sizeof(st->lv)); memset(st->la, 0, sizeof(st->la));
  memset(st->lf, 0, sizeof(st->lf)); memset(st->lu_passive, 0, sizeof(st->lu_passive)); memset(st->lu, 0, sizeof(st->lu));
  memset(st->Fq, 0, sizeof(st->Fq)); memset(st->Fv, 0, sizeof(st->Fv)); memset(st->P, 0, sizeof(st->P));
}
static void zero_matrix(stage_t* st) {
  memset(st->Qxx, 0, sizeof(st->Qxx)); memset(st->Qxu, 0, sizeof(st->Qxu)); memset(st->Quu, 0, sizeof(st->Quu));
  memset(st->Qaa, 0, sizeof(st->Qaa)); memset(st->Qff, 0, sizeof(st->Qff));
  memset(st->Fvq, 0, sizeof(st->Fvq)); memset(st->Fvv, 0, sizeof(st->Fvv)); memset(st->Fvu, 0, sizeof(st->Fvu));
  memset(st->Fqq6, 0, sizeof(st->Fqq6)); memset(st->Fqv6, 0, sizeof(st->Fqv6)); memset(st->Pq, 0, sizeof(st->Pq));
}
/* linearizeForwardEuler / linearizeImpulseForwardEuler */
static void linearize_state_equation(stage_t* st, int impulse, double dt, const double* q_prev, const next_t* nx) {
  fb_subtract(st->q, nx->q, st->Fq);
  if (!impulse) {
    for (int j = 0; j < NV; ++j) st->Fq[j] = fma(dt, st->v[j], st->Fq[j]);
    for (int j = 0; j < NV; ++j) st->Fv[j] = fma(dt, st->a[j], st->v[j]) - nx->v[j];
  } else {
    for (int j = 0; j < NV; ++j) st->Fv[j] = (st->v[j] + st->a[j]) - nx->v[j];
  }
  fb_dsubtract_dplus(st->q, nx->q, st->Fqq6);
  fb_dsubtract_dminus(q_prev, st->q, st->Fqq_prev6);
  mv(MM_ADD, 6, 6, st->Fqq6, 1, 6, nx->lmd, st->lq);
  mv(MM_ADD, 6, 6, st->Fqq_prev6, 1, 6, st->lmd, st->lq);
  for (int j = 6; j < NV; ++j) st->lq[j] += nx->lmd[j] - st->lmd[j];
  if (!impulse) {
    for (int j = 0; j < NV; ++j) st->lv[j] += (fma(dt, nx->lmd[j], nx->gmm[j]) - st->gmm[j]);
    for (int j = 0; j < NV; ++j) st->la[j] = fma(dt, nx->gmm[j], st->la[j]);
  } else {
    for (int j = 0; j < NV; ++j) st->lv[j] += (nx->gmm[j] - st->gmm[j]);
    for (int j = 0; j < NV; ++j) st->la[j] += nx->gmm[j];
  }
}
/* condenseForwardEuler / condenseImpulseForwardEuler */
static void condense_state_equation(stage_t* st, int impulse, double dt, const double* q_next) {
  double tmp6[36], fq[6];
  fb_dsubtract_inverse(st->Fqq_prev6, st->Fqq_prev_inv);
  fb_dsubtract_dminus(st->q, q_next, tmp6);
  fb_dsubtract_inverse(tmp6, st->Fqq_inv);
  memcpy(st->Fqq_prev6, st->Fqq6, sizeof(tmp6));
  for (int i = 0; i < 6; ++i) { st->Fq_prev[i] = st->Fq[i]; fq[i] = st->Fq[i]; }
  mm(MM_SET, 6, 6, 6, st->Fqq_inv, 6, 1, st->Fqq_prev6, 6, 1, st->Fqq6, 6);
  for (int i = 0; i < 36; ++i) st->Fqq6[i] = -st->Fqq6[i];
  if (!impulse)
    for (int i = 0; i < 36; ++i) st->Fqv6[i] = -dt * st->Fqq_inv[i];
  mv(MM_SET, 6, 6, st->Fqq_inv, 6, 1, fq, st->Fq);
  for (int i = 0; i < 6; ++i) st->Fq[i] = -st->Fq[i];
}

/* ---------------------------------------------------------------------------------------------- */
/* ContactDynamics (ocp/contact_dynamics.hxx) and ImpulseDynamicsForwardEuler                        */
/* ---------------------------------------------------------------------------------------------- */
static void stack_active(const stage_t* st, const double x[FB_NC][3], double* out) {
  int k = 0;
  for (int i = 0; i < FB_NC; ++i)
    if (st->active[i]) { out[3 * k] = x[i][0]; out[3 * k + 1] = x[i][1]; out[3 * k + 2] = x[i][2]; ++k; }
}
static void masked_forces(const stage_t* st, double f[FB_NC][3]) {
  for (int i = 0; i < FB_NC; ++i)
    for (int x = 0; x < 3; ++x) f[i][x] = st->active[i] ? st->f[i][x] : 0.0;
}

/* ContactDistance on a stage whose kinematics `kin` are up to date (contact_distance.cpp:73-110,134-150): for every contact
 * that is NOT active, residual = -z + slack with z the height of the contact frame, duality, and the dual residual
 * lq -= dt dual J2 with J2 = row 2 of getFrameJacobian (the LOCAL-frame Jacobian -- the reference differentiates the world
 * height with the local z row; kept).  Evaluated after the dynamics terms because it needs the same kinematics. */
static void contact_distance_stage(const oracle_fb_problem_t* p, stage_t* st, const fb_kin_t* kin, double dt) {
  cdata_t* d = &st->c[C_DISTANCE];
  for (int i = 0; i < FB_NC; ++i) {
    d->residual[i] = 0.0;
    d->duality[i] = 0.0;
    if (st->active[i]) continue;
    fb_frame_t fr;
    fb_frame_kinematics(kin, i, 0, &fr);
    st->cdz[i] = fr.P[2];
    for (int c = 0; c < NV; ++c) st->cdJ[i][c] = fr.J[2][c];
    if (p->enable_distance == 2) {   /* consistent variant: row 2 of the WORLD-aligned Jacobian R_f J_lin = d z / d q */
      const double* Rf = kin->R[1 + ANYMAL_CONTACT_PARENT_JOINT[i]];
      for (int c = 0; c < NV; ++c) st->cdJ[i][c] = fma(Rf[8], fr.J[2][c], fma(Rf[7], fr.J[1][c], Rf[6] * fr.J[0][c]));
    }
    d->residual[i] = -fr.P[2] + d->slack[i];
    d->duality[i] = d->slack[i] * d->dual[i] - p->barrier;
    for (int c = 0; c < NV; ++c) st->lq[c] -= (dt * d->dual[i]) * st->cdJ[i][c];
  }
}
/* ContactDistance::condenseSlackAndDual (contact_distance.cpp:87-110): lq -= (dt (dual residual - duality) / slack) J2 before
 * the contact dynamics are condensed (part 0), Qqq += (dt dual / slack) J2^T J2 (part 1).  Part 1 -- the one dense term of the
 * stage Hessian -- is added AFTER the condensed products: a summation order chosen so that the product kernels of the GPU path
 * are the same with and without this component. */
static void condense_contact_distance(stage_t* st, double dt, int part) {
  const cdata_t* d = &st->c[C_DISTANCE];
  for (int i = 0; i < FB_NC; ++i) {
    if (st->active[i]) continue;
    const double rs = 1.0 / d->slack[i];
    const double w = (dt * d->dual[i]) * rs;
    const double g = (dt * fma(d->dual[i], d->residual[i], -d->duality[i])) * rs;
    for (int r = 0; r < NV; ++r) {
      if (part == 0) { st->lq[r] -= g * st->cdJ[i][r]; continue; }
      const double wr = w * st->cdJ[i][r];
      for (int c = 0; c < NV; ++c) st->Qxx[r * NX + c] += wr * st->cdJ[i][c];
    }
  }
}

/* linearizeContactDynamics (:48-83) resp. linearizeImpulseDynamics (impulse_dynamics_forward_euler.hxx:25-45);
 * residual_only: computeContactDynamicsResidual-style evaluation is a subset (derivatives skipped by the caller). */
static void linearize_contact_dynamics(const oracle_fb_problem_t* p, stage_t* st, int impulse, double dt) {
  const int dimf = st->dimf;
  const double baumgarte = p->T / p->N;
  double f[FB_NC][3], mu_stack[MAXF], dq[NV * NV], dv[NV * NV];
  fb_kin_t kin;
  masked_forces(st, f);
  stack_active(st, st->mu, mu_stack);
  memset(st->dIDCdqv, 0, sizeof(st->dIDCdqv));
  memset(st->dCda, 0, sizeof(st->dCda));
  memset(st->IDC, 0, sizeof(st->IDC));
  if (!impulse) {
    fb_forward_kinematics(st->q, st->v, st->a, &kin);
    fb_rnea_derivatives(&kin, f, ANYMAL_GRAVITY, st->IDC, dq, dv, st->Mm);
    for (int j = 0; j < NU; ++j) st->IDC[6 + j] -= st->u[j];
  } else {
    fb_forward_kinematics(st->q, NULL, st->a, &kin);                  /* RNEAImpulse: zero velocity, a = dv, no gravity */
    fb_rnea_derivatives(&kin, f, 0.0, st->IDC, dq, NULL, st->Mm);
    memset(dv, 0, sizeof(dv));
    double vpdv[NV];
    for (int j = 0; j < NV; ++j) vpdv[j] = st->v[j] + st->a[j];
    fb_forward_kinematics(st->q, vpdv, NULL, &kin);                   /* updateKinematics(q, v + dv) */
  }
  for (int r = 0; r < NV; ++r)
    for (int c = 0; c < NV; ++c) { st->dIDCdqv[r * NX + c] = dq[r * NV + c]; st->dIDCdqv[r * NX + NV + c] = dv[r * NV + c]; }
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    if (!st->active[i]) continue;
    fb_frame_t fr;
    if (!impulse) {
      double dCdq[3 * NV], dCdv[3 * NV], dCda[3 * NV];
      fb_frame_kinematics(&kin, i, 2, &fr);
      fb_baumgarte_residual(&fr, baumgarte, st->cpoints[i], st->IDC + NV + 3 * k);
      fb_baumgarte_derivatives(&kin, i, &fr, baumgarte, dCdq, dCdv, dCda);
      for (int x = 0; x < 3; ++x)
        for (int c = 0; c < NV; ++c) {
          st->dIDCdqv[(NV + 3 * k + x) * NX + c] = dCdq[x * NV + c];
          st->dIDCdqv[(NV + 3 * k + x) * NX + NV + c] = dCdv[x * NV + c];
          st->dCda[(3 * k + x) * NV + c] = dCda[x * NV + c];
        }
    } else {
      /* computeContactVelocityResidual / Derivatives (point_contact.hxx:147-176): C = v_lin, dC/dq, dC/dv = dC/ddv = J_lin */
      fb_frame_kinematics(&kin, i, 1, &fr);
      for (int x = 0; x < 3; ++x) {
        st->IDC[NV + 3 * k + x] = fr.vF[x];
        for (int c = 0; c < NV; ++c) {
          st->dIDCdqv[(NV + 3 * k + x) * NX + c] = fr.v_dq[x][c];
          st->dIDCdqv[(NV + 3 * k + x) * NX + NV + c] = fr.J[x][c];
          st->dCda[(3 * k + x) * NV + c] = fr.J[x][c];
        }
      }
    }
    ++k;
  }
  /* augment the (impulse) inverse dynamics constraint */
  const double* dIDdq = st->dIDCdqv;
  const double* dIDdv = st->dIDCdqv + NV;
  const double* dCdq = st->dIDCdqv + NV * NX;
  const double* dCdv = st->dIDCdqv + NV * NX + NV;
  double t18[NV], tf[MAXF];
  mv(MM_SET, NV, NV, dIDdq, 1, NX, st->beta, t18);
  for (int j = 0; j < NV; ++j) st->lq[j] = fma(dt, t18[j], st->lq[j]);
  if (!impulse) {
    mv(MM_SET, NV, NV, dIDdv, 1, NX, st->beta, t18);
    for (int j = 0; j < NV; ++j) st->lv[j] = fma(dt, t18[j], st->lv[j]);
  }
  mv(MM_SET, NV, NV, st->Mm, 1, NV, st->beta, t18);
  for (int j = 0; j < NV; ++j) st->la[j] = fma(dt, t18[j], st->la[j]);
  if (dimf > 0) {
    mv(MM_SET, dimf, NV, st->dCda, NV, 1, st->beta, tf);
    for (int j = 0; j < dimf; ++j) st->lf[j] = fma(-dt, tf[j], st->lf[j]);
  }
  if (!impulse) {
    for (int j = 0; j < NPASS; ++j) st->lu_passive[j] = fma(-dt, st->beta[j], dt * st->nu_passive[j]);
    for (int j = 0; j < NU; ++j) st->lu[j] = fma(-dt, st->beta[6 + j], st->lu[j]);
  }
  if (dimf > 0) {
    mv(MM_SET, NV, dimf, dCdq, 1, NX, mu_stack, t18);
    for (int j = 0; j < NV; ++j) st->lq[j] = fma(dt, t18[j], st->lq[j]);
    mv(MM_SET, NV, dimf, dCdv, 1, NX, mu_stack, t18);
    for (int j = 0; j < NV; ++j) st->lv[j] = fma(dt, t18[j], st->lv[j]);
    mv(MM_SET, NV, dimf, st->dCda, 1, NV, mu_stack, t18);
    for (int j = 0; j < NV; ++j) st->la[j] = fma(dt, t18[j], st->la[j]);
  }
  if (!impulse && st->cactive[C_DISTANCE]) contact_distance_stage(p, st, &kin, dt);
}

/* condenseContactDynamics (:105-158) / condenseImpulseDynamics (impulse_dynamics_forward_euler.hxx:64-105) */
static void condense_contact_dynamics(stage_t* st, int impulse, double dt) {
  const int dimf = st->dimf, nvf = NV + dimf;
  const int info = fb_MJtJinv(st->Mm, st->dCda, dimf, st->MJtJinv, NVF);
  if (info && !st->chol_info) st->chol_info = info;
  mm(MM_SET, nvf, NX, nvf, st->MJtJinv, NVF, 1, st->dIDCdqv, NX, 1, st->MJ_dIDC, NX);
  mv(MM_SET, nvf, nvf, st->MJtJinv, NVF, 1, st->IDC, st->MJ_IDC);
  for (int r = 0; r < NV; ++r)
    for (int c = 0; c < NX; ++c) st->Qafqv[r * NX + c] = -st->Qaa[r] * st->MJ_dIDC[r * NX + c];
  mm(MM_SET, dimf, NX, dimf, st->Qff, MAXF, 1, st->MJ_dIDC + NV * NX, NX, 1, st->Qafqv + NV * NX, NX);
  for (int r = 0; r < dimf; ++r)
    for (int c = 0; c < NX; ++c) st->Qafqv[(NV + r) * NX + c] = -st->Qafqv[(NV + r) * NX + c];
  if (!impulse) {
    for (int r = 0; r < NV; ++r)
      for (int c = 0; c < NV; ++c) st->Qafu[r * NV + c] = st->Qaa[r] * st->MJtJinv[r * NVF + c];
    mm(MM_SET, dimf, NV, dimf, st->Qff, MAXF, 1, st->MJtJinv + NV * NVF, NVF, 1, st->Qafu + NV * NV, NV);
  }
  for (int j = 0; j < NV; ++j) st->laf[j] = fma(-st->Qaa[j], st->MJ_IDC[j], st->la[j]);
  for (int j = 0; j < dimf; ++j) st->laf[NV + j] = -st->lf[j];
  mv(MM_SUB, dimf, dimf, st->Qff, MAXF, 1, st->MJ_IDC + NV, st->laf + NV);
  mm(MM_SUB, NX, NX, nvf, st->MJ_dIDC, 1, NX, st->Qafqv, NX, 1, st->Qxx, NX);
  mv(MM_SUB, NV, nvf, st->MJ_dIDC, 1, NX, st->laf, st->lq);
  mv(MM_SUB, NV, nvf, st->MJ_dIDC + NV, 1, NX, st->laf, st->lv);
  if (!impulse) {
    mm(MM_SUB, NX, NV, nvf, st->MJ_dIDC, 1, NX, st->Qafu, NV, 1, st->Qxu, NV);
    mm(MM_ADD, NV, NV, nvf, st->MJtJinv, NVF, 1, st->Qafu, NV, 1, st->Quu, NV);
    mv(MM_ADD, NPASS, nvf, st->MJtJinv, NVF, 1, st->laf, st->lu_passive);
    mv(MM_ADD, NU, nvf, st->MJtJinv + NPASS * NVF, NVF, 1, st->laf, st->lu);
    for (int r = 0; r < NV; ++r)
      for (int c = 0; c < NV; ++c) {
        st->Fvq[r * NV + c] = -dt * st->MJ_dIDC[r * NX + c];
        st->Fvv[r * NV + c] = -dt * st->MJ_dIDC[r * NX + NV + c] + (r == c ? 1.0 : 0.0);
      }
    for (int r = 0; r < NV; ++r)
      for (int c = 0; c < NU; ++c) st->Fvu[r * NU + c] = dt * st->MJtJinv[r * NVF + NPASS + c];
    for (int j = 0; j < NV; ++j) st->Fv[j] = fma(-dt, st->MJ_IDC[j], st->Fv[j]);
  } else {
    for (int r = 0; r < NV; ++r)
      for (int c = 0; c < NV; ++c) {
        st->Fvq[r * NV + c] = -st->MJ_dIDC[r * NX + c];
        st->Fvv[r * NV + c] = (r == c ? 1.0 : 0.0) - st->MJ_dIDC[r * NX + NV + c];
      }
    for (int j = 0; j < NV; ++j) st->Fv[j] -= st->MJ_IDC[j];
  }
}

/* ForwardSwitchingConstraint::linearizeSwitchingConstraint (forward_switching_constraint.hxx:27-68) */
static void switching_residual(stage_t* st, double dt1, double dt2, double* dq_out, fb_kin_t* kin) {
  double dqv[NV], q2[NQ];
  const double c1 = dt1 + dt2, c2 = dt1 * dt2;
  for (int j = 0; j < NV; ++j) dqv[j] = fma(c2, st->a[j], c1 * st->v[j]);
  fb_integrate(st->q, dqv, 1.0, q2);
  fb_forward_kinematics(q2, NULL, NULL, kin);
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    if (!st->imp_active[i]) continue;
    double P[3];
    fb_contact_point(kin, i, P);
    for (int x = 0; x < 3; ++x) st->P[3 * k + x] = P[x] - st->ipoints[i][x];
    ++k;
  }
  if (dq_out) memcpy(dq_out, dqv, sizeof(dqv));
}
static void linearize_switching_constraint(stage_t* st, double dt1, double dt2) {
  const int dimi = st->dimi;
  double dqv[NV], Jq6[36], Jv6[36];
  fb_kin_t kin;
  switching_residual(st, dt1, dt2, dqv, &kin);
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    if (!st->imp_active[i]) continue;
    fb_frame_t fr;
    fb_frame_kinematics(&kin, i, 0, &fr);
    const double* Rf = kin.R[1 + ANYMAL_CONTACT_PARENT_JOINT[i]];
    for (int c = 0; c < NV; ++c) {
      const double Jl[3] = {fr.J[0][c], fr.J[1][c], fr.J[2][c]};
      double w[3];
      fb_rot(Rf, Jl, w);
      for (int x = 0; x < 3; ++x) st->Pq[(3 * k + x) * NV + c] = w[x];
    }
    ++k;
  }
  fb_dintegrate_dq(dqv, Jq6);
  fb_dintegrate_dv(dqv, Jv6);
  const double c1 = dt1 + dt2, c2 = dt1 * dt2;
  /* Phiq = Pq dintegrate_dq ; Phiv = (dt1+dt2) Pq dintegrate_dv ; Phia = dt1 dt2 Pq dintegrate_dv */
  double PJv[MAXF * NV];
  mm(MM_SET, dimi, 6, 6, st->Pq, NV, 1, Jq6, 6, 1, st->Phix, NX);
  mm(MM_SET, dimi, 6, 6, st->Pq, NV, 1, Jv6, 6, 1, PJv, NV);
  for (int r = 0; r < dimi; ++r) {
    for (int c = 6; c < NV; ++c) { st->Phix[r * NX + c] = st->Pq[r * NV + c]; PJv[r * NV + c] = st->Pq[r * NV + c]; }
    for (int c = 0; c < NV; ++c) {
      st->Phix[r * NX + NV + c] = c1 * PJv[r * NV + c];
      st->Phia[r * NV + c] = c2 * PJv[r * NV + c];
    }
  }
  mv(MM_ADD, NV, dimi, st->Phix, 1, NX, st->xi, st->lq);
  mv(MM_ADD, NV, dimi, st->Phix + NV, 1, NX, st->xi, st->lv);
  mv(MM_ADD, NV, dimi, st->Phia, 1, NV, st->xi, st->la);
}
/* ContactDynamics::condenseSwitchingConstraint (contact_dynamics.hxx:194-200) */
static void condense_switching_constraint(stage_t* st) {
  const int dimi = st->dimi;
  mm(MM_SUB, dimi, NX, NV, st->Phia, NV, 1, st->MJ_dIDC, NX, 1, st->Phix, NX);
  mm(MM_SET, dimi, NU, NV, st->Phia, NV, 1, st->MJtJinv + NPASS, NVF, 1, st->Phiu, NU);
  mv(MM_SUB, dimi, NV, st->Phia, NV, 1, st->MJ_IDC, st->P);
}

/* SplitOCP::linearizeOCP / computeKKTResidual (split_ocp.hxx:58-134,196-260), ImpulseSplitOCP twins */
static void linearize_stage(const oracle_fb_problem_t* p, stage_t* st, const elem_t* e, const double* q_prev, const next_t* nx,
                            int residual_only) {
  const int impulse = e->kind == K_IMPULSE;
  const double dt = impulse ? 1.0 : e->dt;
  st->chol_info = 0;
  zero_residual(st);
  if (!residual_only) zero_matrix(st);
  cost_derivatives(p, st, e->kind, dt, 0);
  if (residual_only) primal_dual_residual(p, st);
  augment_dual_residual(p, st, dt);
  linearize_state_equation(st, impulse, dt, q_prev, nx);
  if (!residual_only) condense_state_equation(st, impulse, dt, nx->q);
  linearize_contact_dynamics(p, st, impulse, dt);
  if (residual_only) {
    if (e->sw_impulse >= 0) linearize_switching_constraint(st, e->dt, e->dt_next);
    return;
  }
  cost_derivatives(p, st, e->kind, dt, 2);   /* Hessian part only, see below */
  condense_slack_and_dual(p, st, dt);
  if (!impulse && st->cactive[C_DISTANCE]) condense_contact_distance(st, dt, 0);
  if (e->sw_impulse >= 0) linearize_switching_constraint(st, e->dt, e->dt_next);
  condense_contact_dynamics(st, impulse, dt);
  if (e->sw_impulse >= 0) condense_switching_constraint(st);
  if (!impulse && st->cactive[C_DISTANCE]) condense_contact_distance(st, dt, 1);
}

/* TerminalOCP::linearizeOCP / computeKKTResidual (terminal_ocp.hxx:50-66,120-133) */
static void linearize_terminal(const oracle_fb_problem_t* p, stage_t* st, const double* q_prev, int residual_only) {
  memset(st->lq, 0, sizeof(st->lq));
  memset(st->lv, 0, sizeof(st->lv));
  cost_derivatives(p, st, K_TERMINAL, 1.0, 0);
  fb_dsubtract_dminus(q_prev, st->q, st->Fqq_prev6);
  mv(MM_ADD, 6, 6, st->Fqq_prev6, 1, 6, st->lmd, st->lq);
  for (int j = 6; j < NV; ++j) st->lq[j] -= st->lmd[j];
  for (int j = 0; j < NV; ++j) st->lv[j] -= st->gmm[j];
  if (residual_only) return;
  fb_dsubtract_inverse(st->Fqq_prev6, st->Fqq_prev_inv);
  memset(st->Qxx, 0, sizeof(st->Qxx));
  cost_derivatives(p, st, K_TERMINAL, 1.0, 2);
}

/* ---------------------------------------------------------------------------------------------- */
/* Riccati recursion                                                                               */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { const double *Pqq, *Pqv, *Pvv, *sq, *sv; } ric_t;

/* BackwardRiccatiRecursionFactorizer::factorizeKKTMatrix (backward_riccati_recursion_factorizer.hxx:44-111) and the
 * impulse twin (impulse_backward_riccati_recursion_factorizer.hxx:33-70).  AtP* are kept for the s recursion. */
typedef struct { double AtPqq[NV * NV], AtPqv[NV * NV], AtPvq[NV * NV], AtPvv[NV * NV], BtPq[NU * NV], BtPv[NU * NV]; } ricwork_t;

static void factorize_kkt_matrix(stage_t* st, int impulse, double dt, const ric_t* rn, ricwork_t* w) {
  double* Qqq = st->Qxx;
  double* Qqv = st->Qxx + NV;
  double* Qvv = st->Qxx + NV * NX + NV;
  /* A^T P, top 6 rows through the 6x6 SE(3) blocks, the rest identity / dt */
  mm(MM_SET, 6, NV, 6, st->Fqq6, 1, 6, rn->Pqq, NV, 1, w->AtPqq, NV);
  mm(MM_SET, 6, NV, 6, st->Fqq6, 1, 6, rn->Pqv, NV, 1, w->AtPqv, NV);
  for (int r = 6; r < NV; ++r)
    for (int c = 0; c < NV; ++c) { w->AtPqq[r * NV + c] = rn->Pqq[r * NV + c]; w->AtPqv[r * NV + c] = rn->Pqv[r * NV + c]; }
  if (!impulse) {
    mm(MM_SET, 6, NV, 6, st->Fqv6, 1, 6, rn->Pqq, NV, 1, w->AtPvq, NV);
    mm(MM_SET, 6, NV, 6, st->Fqv6, 1, 6, rn->Pqv, NV, 1, w->AtPvv, NV);
    for (int r = 6; r < NV; ++r)
      for (int c = 0; c < NV; ++c) { w->AtPvq[r * NV + c] = dt * rn->Pqq[r * NV + c]; w->AtPvv[r * NV + c] = dt * rn->Pqv[r * NV + c]; }
  }
  /* += Fvq^T Pqv^T etc. (Pvq = Pqv^T) */
  mm(MM_ADD, NV, NV, NV, st->Fvq, 1, NV, rn->Pqv, 1, NV, w->AtPqq, NV);
  mm(MM_ADD, NV, NV, NV, st->Fvq, 1, NV, rn->Pvv, NV, 1, w->AtPqv, NV);
  if (!impulse) {
    mm(MM_ADD, NV, NV, NV, st->Fvv, 1, NV, rn->Pqv, 1, NV, w->AtPvq, NV);
    mm(MM_ADD, NV, NV, NV, st->Fvv, 1, NV, rn->Pvv, NV, 1, w->AtPvv, NV);
    mm(MM_SET, NU, NV, NV, st->Fvu, 1, NU, rn->Pqv, 1, NV, w->BtPq, NV);
    mm(MM_SET, NU, NV, NV, st->Fvu, 1, NU, rn->Pvv, NV, 1, w->BtPv, NV);
  } else {
    mm(MM_SET, NV, NV, NV, st->Fvv, 1, NV, rn->Pqv, 1, NV, w->AtPvq, NV);
    mm(MM_SET, NV, NV, NV, st->Fvv, 1, NV, rn->Pvv, NV, 1, w->AtPvv, NV);
  }
  /* Factorize F */
  mm(MM_ADD, NV, 6, 6, w->AtPqq, NV, 1, st->Fqq6, 6, 1, Qqq, NX);
  for (int r = 0; r < NV; ++r)
    for (int c = 6; c < NV; ++c) Qqq[r * NX + c] += w->AtPqq[r * NV + c];
  if (!impulse) {
    mm(MM_ADD, NV, 6, 6, w->AtPqq, NV, 1, st->Fqv6, 6, 1, Qqv, NX);
    for (int r = 0; r < NV; ++r)
      for (int c = 6; c < NV; ++c) Qqv[r * NX + c] = fma(dt, w->AtPqq[r * NV + c], Qqv[r * NX + c]);
    mm(MM_ADD, NV, 6, 6, w->AtPvq, NV, 1, st->Fqv6, 6, 1, Qvv, NX);
    for (int r = 0; r < NV; ++r)
      for (int c = 6; c < NV; ++c) Qvv[r * NX + c] = fma(dt, w->AtPvq[r * NV + c], Qvv[r * NX + c]);
  }
  mm(MM_ADD, NV, NV, NV, w->AtPqv, NV, 1, st->Fvq, NV, 1, Qqq, NX);
  mm(MM_ADD, NV, NV, NV, w->AtPqv, NV, 1, st->Fvv, NV, 1, Qqv, NX);
  mm(MM_ADD, NV, NV, NV, w->AtPvv, NV, 1, st->Fvv, NV, 1, Qvv, NX);
  if (impulse) return;
  /* Factorize H, G and the vector term: Qqu, Qvu are the actuated columns (6..17) of Qxu */
  mm(MM_ADD, NV, NU, NV, w->AtPqv, NV, 1, st->Fvu, NU, 1, st->Qxu + NPASS, NV);
  mm(MM_ADD, NV, NU, NV, w->AtPvv, NV, 1, st->Fvu, NU, 1, st->Qxu + NV * NV + NPASS, NV);
  mm(MM_ADD, NU, NU, NV, w->BtPv, NV, 1, st->Fvu, NU, 1, st->Quu + NPASS * NV + NPASS, NV);
  mv(MM_ADD, NU, NV, w->BtPq, NV, 1, st->Fq, st->lu);
  mv(MM_ADD, NU, NV, w->BtPv, NV, 1, st->Fv, st->lu);
  mv(MM_SUB, NU, NV, st->Fvu, 1, NU, rn->sv, st->lu);
}

/* factorizeRiccatiFactorization (:114-161) and the impulse twin (:73-102) */
static void factorize_riccati(stage_t* st, int impulse, double dt, const ric_t* rn, const ricwork_t* w) {
  const double* Qqq = st->Qxx;
  const double* Qqv = st->Qxx + NV;
  const double* Qvv = st->Qxx + NV * NX + NV;
  for (int r = 0; r < NV; ++r)
    for (int c = 0; c < NV; ++c) {
      st->Pqq[r * NV + c] = Qqq[r * NX + c];
      st->Pqv[r * NV + c] = Qqv[r * NX + c];
      st->Pvv[r * NV + c] = Qvv[r * NX + c];
    }
  if (!impulse) {
    double GK[NU * NX];
    mm(MM_SET, NU, NX, NU, st->Quu + NPASS * NV + NPASS, NV, 1, st->K, NX, 1, GK, NX);
    mm(MM_SUB, NV, NV, NU, st->K, 1, NX, GK, NX, 1, st->Pqq, NV);
    mm(MM_SUB, NV, NV, NU, st->K, 1, NX, GK + NV, NX, 1, st->Pqv, NV);
    mm(MM_SUB, NV, NV, NU, st->K + NV, 1, NX, GK + NV, NX, 1, st->Pvv, NV);
  }
  /* preserve the symmetry */
  for (int r = 0; r < NV; ++r)
    for (int c = r; c < NV; ++c) {
      const double a = 0.5 * (st->Pqq[r * NV + c] + st->Pqq[c * NV + r]);
      const double b = 0.5 * (st->Pvv[r * NV + c] + st->Pvv[c * NV + r]);
      st->Pqq[r * NV + c] = a; st->Pqq[c * NV + r] = a;
      st->Pvv[r * NV + c] = b; st->Pvv[c * NV + r] = b;
    }
  mv(MM_SET, 6, 6, st->Fqq6, 1, 6, rn->sq, st->sq);
  for (int j = 6; j < NV; ++j) st->sq[j] = rn->sq[j];
  if (!impulse) {
    mv(MM_SET, 6, 6, st->Fqv6, 1, 6, rn->sq, st->sv);
    for (int j = 6; j < NV; ++j) st->sv[j] = dt * rn->sq[j];
    mv(MM_ADD, NV, NV, st->Fvq, 1, NV, rn->sv, st->sq);
    mv(MM_ADD, NV, NV, st->Fvv, 1, NV, rn->sv, st->sv);
  } else {
    mv(MM_ADD, NV, NV, st->Fvq, 1, NV, rn->sv, st->sq);
    mv(MM_SET, NV, NV, st->Fvv, 1, NV, rn->sv, st->sv);
  }
  mv(MM_SUB, NV, NV, w->AtPqq, NV, 1, st->Fq, st->sq);
  mv(MM_SUB, NV, NV, w->AtPqv, NV, 1, st->Fv, st->sq);
  mv(MM_SUB, NV, NV, w->AtPvq, NV, 1, st->Fq, st->sv);
  mv(MM_SUB, NV, NV, w->AtPvv, NV, 1, st->Fv, st->sv);
  for (int j = 0; j < NV; ++j) { st->sq[j] -= st->lq[j]; st->sv[j] -= st->lv[j]; }
  if (!impulse) {
    mv(MM_SUB, NV, NU, st->Qxu + NPASS, NV, 1, st->k, st->sq);
    mv(MM_SUB, NV, NU, st->Qxu + NV * NV + NPASS, NV, 1, st->k, st->sv);
  }
}

/* SplitRiccatiFactorizer::backwardRiccatiRecursion, plain (:36-52) and constrained (:55-100);
 * ImpulseSplitRiccatiFactorizer::backwardRiccatiRecursion (impulse_split_riccati_factorizer.hxx:26-33). */
static void riccati_backward_stage(stage_t* st, int impulse, double dt, const ric_t* rn, int constrained) {
  ricwork_t w;
  factorize_kkt_matrix(st, impulse, dt, rn, &w);
  if (impulse) {
    factorize_riccati(st, 1, dt, rn, &w);
    return;
  }
  double G[NU * NU], L[NU * NU], rd[NU];
  for (int r = 0; r < NU; ++r)
    for (int c = 0; c < NU; ++c) G[r * NU + c] = st->Quu[(NPASS + r) * NV + NPASS + c];
  const int info = fb_llt(G, NU, NU, L, NU, rd);
  if (info && !st->chol_info) st->chol_info = 200 + info;
  const double* Qxu = st->Qxu + NPASS;      /* 36 x 12 block, leading dimension NV */
  if (!constrained) {
    /* K = -G^-1 Qxu^T, k = -G^-1 lu */
    for (int c = 0; c < NX; ++c) {
      double col[NU];
      for (int r = 0; r < NU; ++r) col[r] = Qxu[c * NV + r];
      fb_llt_solve(L, NU, rd, NU, col, 1);
      for (int r = 0; r < NU; ++r) st->K[r * NX + c] = -col[r];
    }
    double col[NU];
    for (int r = 0; r < NU; ++r) col[r] = st->lu[r];
    fb_llt_solve(L, NU, rd, NU, col, 1);
    for (int r = 0; r < NU; ++r) st->k[r] = -col[r];
    factorize_riccati(st, 0, dt, rn, &w);
    return;
  }
  const int dimi = st->dimi;
  double Ginv[NU * NU], DGinv[MAXF * NU], S[MAXF * MAXF], Ls[MAXF * MAXF], rds[MAXF], SinvDGinv[MAXF * NU];
  for (int c = 0; c < NU; ++c) {
    for (int r = 0; r < NU; ++r) Ginv[r * NU + c] = (r == c) ? 1.0 : 0.0;
    fb_llt_solve(L, NU, rd, NU, Ginv + c, NU);
  }
  /* DGinv^T = G^-1 Phiu^T */
  for (int r = 0; r < dimi; ++r) {
    for (int c = 0; c < NU; ++c) DGinv[r * NU + c] = st->Phiu[r * NU + c];
    fb_llt_solve(L, NU, rd, NU, DGinv + r * NU, 1);
  }
  mm(MM_SET, dimi, dimi, NU, DGinv, NU, 1, st->Phiu, 1, NU, S, MAXF);
  const int i2 = fb_llt(S, MAXF, dimi, Ls, MAXF, rds);
  if (i2 && !st->chol_info) st->chol_info = 300 + i2;
  for (int c = 0; c < NU; ++c) {
    for (int r = 0; r < dimi; ++r) SinvDGinv[r * NU + c] = DGinv[r * NU + c];
    fb_llt_solve(Ls, MAXF, rds, dimi, SinvDGinv + c, NU);
  }
  mm(MM_SUB, NU, NU, dimi, SinvDGinv, 1, NU, DGinv, NU, 1, Ginv, NU);
  mm(MM_SET, NU, NX, NU, Ginv, NU, 1, Qxu, 1, NV, st->K, NX);
  for (int i = 0; i < NU * NX; ++i) st->K[i] = -st->K[i];
  mm(MM_SUB, NU, NX, dimi, SinvDGinv, 1, NU, st->Phix, NX, 1, st->K, NX);
  mv(MM_SET, NU, NU, Ginv, NU, 1, st->lu, st->k);
  for (int i = 0; i < NU; ++i) st->k[i] = -st->k[i];
  mv(MM_SUB, NU, dimi, SinvDGinv, 1, NU, st->P, st->k);
  /* M = S^-1 Phix - SinvDGinv Qxu^T ; m = S^-1 P - SinvDGinv lu */
  for (int c = 0; c < NX; ++c) {
    for (int r = 0; r < dimi; ++r) st->cM[r * NX + c] = st->Phix[r * NX + c];
    fb_llt_solve(Ls, MAXF, rds, dimi, st->cM + c, NX);
  }
  mm(MM_SUB, dimi, NX, NU, SinvDGinv, NU, 1, Qxu, 1, NV, st->cM, NX);
  for (int r = 0; r < dimi; ++r) st->cm[r] = st->P[r];
  fb_llt_solve(Ls, MAXF, rds, dimi, st->cm, 1);
  mv(MM_SUB, dimi, NU, SinvDGinv, NU, 1, st->lu, st->cm);
  factorize_riccati(st, 0, dt, rn, &w);
  double DtM[NU * NX], KtDtM[NX * NX];
  mm(MM_SET, NU, NX, dimi, st->Phiu, 1, NU, st->cM, NX, 1, DtM, NX);
  mm(MM_SET, NX, NX, NU, st->K, 1, NX, DtM, NX, 1, KtDtM, NX);
  for (int r = 0; r < NV; ++r)
    for (int c = 0; c < NV; ++c) {
      st->Pqq[r * NV + c] = (st->Pqq[r * NV + c] - KtDtM[r * NX + c]) - KtDtM[c * NX + r];
      st->Pqv[r * NV + c] = (st->Pqv[r * NV + c] - KtDtM[r * NX + NV + c]) - KtDtM[(NV + c) * NX + r];
      st->Pvv[r * NV + c] = (st->Pvv[r * NV + c] - KtDtM[(NV + r) * NX + NV + c]) - KtDtM[(NV + c) * NX + NV + r];
    }
  mv(MM_SUB, NV, dimi, st->Phix, 1, NX, st->cm, st->sq);
  mv(MM_SUB, NV, dimi, st->Phix + NV, 1, NX, st->cm, st->sv);
}

/* SplitRiccatiFactorizer::forwardRiccatiRecursion (:103-128) / impulse twin (:36-52): d_next.dx from d.dx */
static void riccati_forward_stage(stage_t* st, int impulse, double dt, double* dq_next, double* dv_next) {
  if (!impulse) {
    double dx[NX];
    memcpy(dx, st->dq, sizeof(st->dq));
    memcpy(dx + NV, st->dv, sizeof(st->dv));
    mv(MM_SET, NU, NX, st->K, NX, 1, dx, st->du);
    for (int j = 0; j < NU; ++j) st->du[j] += st->k[j];
  }
  for (int j = 0; j < NV; ++j) { dq_next[j] = st->Fq[j]; dv_next[j] = st->Fv[j]; }
  mv(MM_ADD, 6, 6, st->Fqq6, 6, 1, st->dq, dq_next);
  for (int j = 6; j < NV; ++j) dq_next[j] += st->dq[j];
  if (!impulse) {
    mv(MM_ADD, 6, 6, st->Fqv6, 6, 1, st->dv, dq_next);
    for (int j = 6; j < NV; ++j) dq_next[j] = fma(dt, st->dv[j], dq_next[j]);
  }
  mv(MM_ADD, NV, NV, st->Fvq, NV, 1, st->dq, dv_next);
  mv(MM_ADD, NV, NV, st->Fvv, NV, 1, st->dv, dv_next);
  if (!impulse) mv(MM_ADD, NV, NU, st->Fvu, NU, 1, st->du, dv_next);
}
/* computeCostateDirection (:131-139) */
static void costate_direction(stage_t* st) {
  mv(MM_SET, NV, NV, st->Pqq, NV, 1, st->dq, st->dlmd);
  mv(MM_ADD, NV, NV, st->Pqv, NV, 1, st->dv, st->dlmd);
  mv(MM_SET, NV, NV, st->Pqv, 1, NV, st->dq, st->dgmm);
  mv(MM_ADD, NV, NV, st->Pvv, NV, 1, st->dv, st->dgmm);
  for (int j = 0; j < NV; ++j) { st->dlmd[j] -= st->sq[j]; st->dgmm[j] -= st->sv[j]; }
}
/* ContactDynamics::computeCondensedPrimalDirection (:161-168) / ImpulseDynamicsForwardEuler::expansionPrimal */
static void condensed_primal_direction(stage_t* st, int impulse) {
  const int nvf = NV + st->dimf;
  double dx[NX];
  memcpy(dx, st->dq, sizeof(st->dq));
  memcpy(dx + NV, st->dv, sizeof(st->dv));
  mv(MM_SET, nvf, NX, st->MJ_dIDC, NX, 1, dx, st->daf);
  for (int j = 0; j < nvf; ++j) st->daf[j] = -st->daf[j];
  if (!impulse) mv(MM_ADD, nvf, NU, st->MJtJinv + NPASS, NVF, 1, st->du, st->daf);
  for (int j = 0; j < nvf; ++j) st->daf[j] -= st->MJ_IDC[j];
  for (int j = NV; j < nvf; ++j) st->daf[j] = -st->daf[j];
}
/* ContactDynamics::computeCondensedDualDirection (:171-190) / expansionDual; then
 * stateequation::correctCostateDirectionForwardEuler (state_equation.hxx:96-108) */
static void condensed_dual_direction(stage_t* st, int impulse, double dt, const double* dgmm_next) {
  const int nvf = NV + st->dimf;
  double dx[NX];
  memcpy(dx, st->dq, sizeof(st->dq));
  memcpy(dx + NV, st->dv, sizeof(st->dv));
  if (!impulse) {
    const double rdt = 1.0 / dt;
    for (int j = 0; j < NPASS; ++j) st->dnu_passive[j] = st->lu_passive[j];
    mv(MM_ADD, NPASS, NU, st->Quu + NPASS, NV, 1, st->du, st->dnu_passive);
    mv(MM_ADD, NPASS, NX, st->Qxu, 1, NV, dx, st->dnu_passive);
    double t6[NPASS];
    mv(MM_SET, NPASS, NV, st->MJtJinv, NVF, 1, dgmm_next, t6);
    for (int j = 0; j < NPASS; ++j) st->dnu_passive[j] = -(fma(dt, t6[j], st->dnu_passive[j])) * rdt;
    mv(MM_ADD, nvf, NX, st->Qafqv, NX, 1, dx, st->laf);
    mv(MM_ADD, nvf, NU, st->Qafu + NPASS, NV, 1, st->du, st->laf);
    for (int j = 0; j < NV; ++j) st->laf[j] = fma(dt, dgmm_next[j], st->laf[j]);
    mv(MM_SET, nvf, nvf, st->MJtJinv, NVF, 1, st->laf, st->dbetamu);
    for (int j = 0; j < nvf; ++j) st->dbetamu[j] = -st->dbetamu[j] * rdt;
  } else {
    mv(MM_ADD, nvf, NX, st->Qafqv, NX, 1, dx, st->laf);
    for (int j = 0; j < NV; ++j) st->laf[j] += dgmm_next[j];
    mv(MM_SET, nvf, nvf, st->MJtJinv, NVF, 1, st->laf, st->dbetamu);
    for (int j = 0; j < nvf; ++j) st->dbetamu[j] = -st->dbetamu[j];
  }
}
static void correct_costate_direction(stage_t* st) {
  double t[6];
  mv(MM_SET, 6, 6, st->Fqq_prev_inv, 1, 6, st->dlmd, t);
  for (int j = 0; j < 6; ++j) { st->Fq_prev[j] = t[j]; st->dlmd[j] = -t[j]; }
}
/* SplitSolution::integrate (split_solution.hxx:215-239) + Constraints::updateSlack / updateDual */
static void update_stage(stage_t* st, int kind, double ap, double ad) {
  double qn[NQ];
  for (int j = 0; j < NV; ++j) { st->lmd[j] = fma(ap, st->dlmd[j], st->lmd[j]); st->gmm[j] = fma(ap, st->dgmm[j], st->gmm[j]); }
  fb_integrate(st->q, st->dq, ap, qn);
  memcpy(st->q, qn, sizeof(qn));
  for (int j = 0; j < NV; ++j) st->v[j] = fma(ap, st->dv[j], st->v[j]);
  if (kind == K_TERMINAL) return;
  for (int j = 0; j < NV; ++j) { st->a[j] = fma(ap, st->daf[j], st->a[j]); st->beta[j] = fma(ap, st->dbetamu[j], st->beta[j]); }
  if (kind != K_IMPULSE) {
    for (int j = 0; j < NU; ++j) st->u[j] = fma(ap, st->du[j], st->u[j]);
    for (int j = 0; j < NPASS; ++j) st->nu_passive[j] = fma(ap, st->dnu_passive[j], st->nu_passive[j]);
  }
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    if (!st->active[i]) continue;
    for (int x = 0; x < 3; ++x) {
      st->f[i][x] = fma(ap, st->daf[NV + 3 * k + x], st->f[i][x]);
      st->mu[i][x] = fma(ap, st->dbetamu[NV + 3 * k + x], st->mu[i][x]);
    }
    ++k;
  }
  if (kind != K_IMPULSE)
    for (int j = 0; j < st->dimi; ++j) st->xi[j] = fma(ap, st->dxi[j], st->xi[j]);
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c]) continue;
    cdata_t* d = &st->c[c];
    for (int j = 0; j < comp_dim(c); ++j) {
      d->slack[j] = fma(ap, d->dslack[j], d->slack[j]);
      d->dual[j] = fma(ad, d->ddual[j], d->dual[j]);
    }
  }
}

/* squaredNormKKTResidual (split_ocp.hxx:263-279, impulse_split_ocp.hxx:131-142, terminal_ocp.hxx:136-142) */
static double stage_kkt_sqnorm(const stage_t* st, int kind, double dt) {
  double e = 0.0;
  if (kind == K_TERMINAL) return sqnorm_n(st->lq, NV) + sqnorm_n(st->lv, NV);
  const int nvf = NV + st->dimf;
  e += sqnorm_n(st->lq, NV) + sqnorm_n(st->lv, NV);
  e += sqnorm_n(st->la, NV);
  e += sqnorm_n(st->lf, st->dimf);
  if (kind != K_IMPULSE) {
    e += sqnorm_n(st->lu_passive, NPASS);
    e += sqnorm_n(st->lu, NU);
  }
  e += sqnorm_n(st->Fq, NV) + sqnorm_n(st->Fv, NV);
  if (kind != K_IMPULSE) {
    e += dt * dt * sqnorm_n(st->IDC, nvf);
    e += dt * dt * constraints_sqnorm(st);
    e += sqnorm_n(st->P, st->dimi);
  } else {
    e += sqnorm_n(st->IDC, nvf);
    e += constraints_sqnorm(st);
  }
  return e;
}

/* ---------------------------------------------------------------------------------------------- */
/* horizon driver                                                                                  */
/* ---------------------------------------------------------------------------------------------- */
static int slot_of(const oracle_fb_ocp_t* o, int kind, int index) {
  const int n1 = o->p.N + 1, m = o->p.max_num_impulse;
  switch (kind) {
    case K_GRID: case K_TERMINAL: return index;
    case K_IMPULSE: return n1 + index;
    case K_AUX: return n1 + m + index;
    default: return n1 + 2 * m + index;
  }
}

/* OCP::discretize + OCPSolver::discretizeSolution (ocp_solver.cpp:283-316): the chain of stages in the order
 * the Riccati recursion visits them, contact / impulse status of every stage. */
static int discretize(oracle_fb_ocp_t* o, double t) {
  oracle_discretization_t* d = &o->disc;
  if (oracle_discretize_ocp(o->cs, o->p.T, o->p.N, t, d) != 0) return -1;
  if (d->N_impulse > o->p.max_num_impulse || d->N_lift > o->p.max_num_impulse) return -2;
  int n = 0;
  for (int i = 0; i < d->N; ++i) {
    elem_t* e = &o->elems[n++];
    *e = (elem_t){K_GRID, i, slot_of(o, K_GRID, i), d->t[i], d->dt[i], d->contact_phase[i], i, -1, 0.0, -1, -1, 0.0};
    if (d->before_impulse_flag[i]) {
      const int k = d->impulse_after[i];
      o->elems[n++] = (elem_t){K_IMPULSE, k, slot_of(o, K_IMPULSE, k), d->t_impulse[k], 0.0, k, -1, -1, 0.0, -1, -1, 0.0};
      o->elems[n++] = (elem_t){K_AUX, k, slot_of(o, K_AUX, k), d->t_impulse[k], d->dt_aux[k], d->contact_phase[i + 1], 0, -1, 0.0, -1, -1, 0.0};
    } else if (d->before_lift_flag[i]) {
      const int k = d->lift_after[i];
      o->elems[n++] = (elem_t){K_LIFT, k, slot_of(o, K_LIFT, k), d->t_lift[k], d->dt_lift[k], d->contact_phase[i + 1], 0, -1, 0.0, -1, -1, 0.0};
    }
  }
  o->elems[n++] = (elem_t){K_TERMINAL, d->N, slot_of(o, K_GRID, d->N), d->t[d->N], 0.0, d->contact_phase[d->N], -1, -1, 0.0, -1, -1, 0.0};
  o->n_elems = n;
  /* switching constraint: imposed two chain elements ahead of an impulse (ocp_linearizer.hxx:139-150,196-214) */
  for (int e = 0; e + 2 < n; ++e)
    if (o->elems[e + 2].kind == K_IMPULSE && o->elems[e].kind != K_AUX && o->elems[e].kind != K_IMPULSE) {
      o->elems[e].sw_impulse = o->elems[e + 2].index;
      o->elems[e].dt_next = o->elems[e + 1].dt;
    }
  /* LineSearch::computeCostAndViolation (line_search.cpp:64-197): a grid stage is always closed with the NEXT GRID
   * stage (the impulse / lift branches of :84-103 are overwritten by the if / else of :104-121) and carries the
   * switching residual whenever that next grid stage precedes an impulse; aux, lift and impulse stages use their
   * chain successor. */
  for (int e = 0; e < n; ++e) {
    elem_t* el = &o->elems[e];
    el->ls_next = e + 1 < n ? e + 1 : -1;
    el->ls_impulse = el->sw_impulse;
    el->ls_dt_next = el->dt_next;
    if (el->kind == K_GRID) {
      int g = e + 1;
      while (g < n && o->elems[g].kind != K_GRID && o->elems[g].kind != K_TERMINAL) ++g;
      el->ls_next = g;
      el->ls_impulse = (g + 1 < n && o->elems[g].kind == K_GRID && o->elems[g + 1].kind == K_IMPULSE) ? o->elems[g + 1].index : -1;
      el->ls_dt_next = o->elems[g].dt;
    }
  }
  /* statuses */
  for (int e = 0; e < n; ++e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    int act[HY_MAX_CONTACTS];
    double pts[HY_MAX_CONTACTS * 3], time;
    if (el->kind == K_IMPULSE) oracle_cs_get_impulse(o->cs, el->phase, act, pts, &time);
    else oracle_cs_get_phase(o->cs, el->phase, act, pts);
    st->dimf = 0;
    for (int i = 0; i < FB_NC; ++i) {
      st->active[i] = act[i];
      st->dimf += 3 * act[i];
      for (int x = 0; x < 3; ++x) st->cpoints[i][x] = pts[3 * i + x];
    }
    st->dimi = 0;
    for (int i = 0; i < FB_NC; ++i) st->imp_active[i] = 0;
    if (el->sw_impulse >= 0) {
      oracle_cs_get_impulse(o->cs, el->sw_impulse, act, pts, &time);
      for (int i = 0; i < FB_NC; ++i) {
        st->imp_active[i] = act[i];
        st->dimi += 3 * act[i];
        for (int x = 0; x < 3; ++x) st->ipoints[i][x] = pts[3 * i + x];
      }
    }
  }
  return d->well_defined ? 0 : -3;
}

static const double* q_prev_of(const oracle_fb_ocp_t* o, int e, const double* q) {
  return e == 0 ? q : o->slots[o->elems[e - 1].slot].q;
}
static void next_of(const oracle_fb_ocp_t* o, int e, next_t* nx) {
  const stage_t* s = &o->slots[o->elems[e + 1].slot];
  nx->lmd = s->lmd; nx->gmm = s->gmm; nx->q = s->q; nx->v = s->v;
}
static void ric_of(const stage_t* s, ric_t* r) { r->Pqq = s->Pqq; r->Pqv = s->Pqv; r->Pvv = s->Pvv; r->sq = s->sq; r->sv = s->sv; }

static void linearize_all(oracle_fb_ocp_t* o, const double* q, int residual_only) {
#pragma omp parallel for schedule(dynamic) num_threads(o->nthreads)
  for (int e = 0; e < o->n_elems; ++e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    if (el->kind == K_TERMINAL) {
      linearize_terminal(&o->p, st, q_prev_of(o, e, q), residual_only);
    } else {
      next_t nx;
      next_of(o, e, &nx);
      linearize_stage(&o->p, st, el, q_prev_of(o, e, q), &nx, residual_only);
    }
    st->kkt_sq = stage_kkt_sqnorm(st, el->kind, el->dt);
  }
}

oracle_fb_ocp_t* oracle_fb_ocp_create(const oracle_fb_problem_t* p, const oracle_contact_sequence_t* cs) {
  oracle_fb_ocp_t* o = (oracle_fb_ocp_t*)calloc(1, sizeof(*o));
  o->p = *p;
  o->cs = cs;
  o->n_slots = p->N + 1 + 3 * p->max_num_impulse;
  o->slots = (stage_t*)calloc((size_t)o->n_slots, sizeof(stage_t));
  o->nthreads = 1;
  for (int s = 0; s < o->n_slots; ++s) o->slots[s].q[6] = 1.0;   /* normalizeConfiguration of the zero vector */
  return o;
}
void oracle_fb_ocp_destroy(oracle_fb_ocp_t* o) {
  if (!o) return;
  free(o->slots);
  free(o);
}
void oracle_fb_ocp_set_threads(oracle_fb_ocp_t* o, int n) { o->nthreads = n > 0 ? n : 1; }
int oracle_fb_problem_size(void) { return (int)sizeof(oracle_fb_problem_t); }

/* OCPSolver::setSolution (ocp_solver.cpp:117-170): broadcast to every stage, impulse, aux and lift slot */
int oracle_fb_ocp_set_solution(oracle_fb_ocp_t* o, const char* name, const double* value) {
  const int n1 = o->p.N + 1, m = o->p.max_num_impulse;
  for (int s = 0; s < o->n_slots; ++s) {
    stage_t* st = &o->slots[s];
    const int is_impulse = s >= n1 && s < n1 + m;
    if (!strcmp(name, "q")) memcpy(st->q, value, sizeof(st->q));
    else if (!strcmp(name, "v")) memcpy(st->v, value, sizeof(st->v));
    else if (!strcmp(name, "a")) memcpy(st->a, value, sizeof(st->a));
    else if (!strcmp(name, "f")) { for (int i = 0; i < FB_NC; ++i) memcpy(st->f[i], value, 3 * sizeof(double)); }
    else if (!strcmp(name, "u")) { if (!is_impulse) memcpy(st->u, value, sizeof(st->u)); }
    else return -1;
  }
  return 0;
}
/* cost reference of one slot (kind, index): q_ref(t), v_ref sampled by the caller at the stage time */
void oracle_fb_ocp_set_reference(oracle_fb_ocp_t* o, int kind, int index, const double* q_ref, const double* v_ref) {
  stage_t* st = &o->slots[slot_of(o, kind, index)];
  memcpy(st->ref_q, q_ref, sizeof(st->ref_q));
  memcpy(st->ref_v, v_ref, sizeof(st->ref_v));
}
/* chain of the current discretisation: rows (kind, index, t, dt, dimf, dimi, cstage) */
int oracle_fb_ocp_discretize(oracle_fb_ocp_t* o, double t) { return discretize(o, t); }
int oracle_fb_ocp_chain(const oracle_fb_ocp_t* o, int* kind, int* index, double* t, double* dt, int* dimf, int* dimi) {
  for (int e = 0; e < o->n_elems; ++e) {
    kind[e] = o->elems[e].kind; index[e] = o->elems[e].index; t[e] = o->elems[e].t; dt[e] = o->elems[e].dt;
    dimf[e] = o->slots[o->elems[e].slot].dimf; dimi[e] = o->slots[o->elems[e].slot].dimi;
  }
  return o->n_elems;
}

/* OCPSolver::initConstraints (ocp_solver.cpp:60-64) -> OCPLinearizer::initConstraints (ocp_linearizer.cpp:40-68):
 * grid stages 0..N_ideal-1 with their index, aux / lift with 0, impulse with -1; the terminal stage has none. */
int oracle_fb_ocp_init_constraints(oracle_fb_ocp_t* o, double t) {
  const int rc = discretize(o, t);
  if (rc == -1 || rc == -2) return rc;
  const int n1 = o->p.N + 1, m = o->p.max_num_impulse;
  for (int s = 0; s < o->n_slots; ++s) {
    stage_t* st = &o->slots[s];
    int ts;
    if (s < o->p.N) ts = s;
    else if (s == o->p.N) { for (int c = 0; c < NCOMP; ++c) st->cactive[c] = 0; continue; }
    else if (s < n1 + m) ts = -1;
    else ts = 0;
    /* every SCHEDULED event, also those beyond the horizon at this t: the reference's discretizer counts
     * contact_sequence.numImpulseEvents() / numLiftEvents() (ocp_discretizer.hxx:246-262) and
     * OCPLinearizer::initConstraints covers all of them (ocp_linearizer.cpp:40-68) */
    int n_phases, n_imp, n_lift;
    oracle_cs_counts(o->cs, &n_phases, &n_imp, &n_lift);
    if (n_imp > m) n_imp = m;
    if (n_lift > m) n_lift = m;
    const int used = (s < n1) || (s < n1 + m ? s - n1 < n_imp
                                  : (s < n1 + 2 * m ? s - n1 - m < n_imp : s - n1 - 2 * m < n_lift));
    set_constraint_stage(&o->p, st, ts);
    if (!used) { for (int c = 0; c < NCOMP; ++c) st->cactive[c] = 0; continue; }
    set_slack_and_dual(&o->p, st);
  }
  return rc;
}


/* ---------------------------------------------------------------------------------------------- */
/* LineSearch (line_search/line_search.hpp:62-158, src/line_search/line_search.cpp:64-197),          */
/* LineSearchFilter (src/line_search/line_search_filter.cpp:34-65)                                   */
/* ---------------------------------------------------------------------------------------------- */
static int filter_is_accepted(const oracle_fb_ocp_t* o, double cost, double viol) {
  for (int i = 0; i < o->filter_n; ++i)
    if (cost >= o->filter_cost[i] && viol >= o->filter_viol[i]) return 0;
  return 1;
}
static void filter_augment(oracle_fb_ocp_t* o, double cost, double viol) {
  int w = 0;
  for (int i = 0; i < o->filter_n; ++i) {
    if (cost <= o->filter_cost[i] && viol <= o->filter_viol[i]) continue;
    o->filter_cost[w] = o->filter_cost[i]; o->filter_viol[w] = o->filter_viol[i]; ++w;
  }
  o->filter_n = w;
  if (w < 256) {
    o->filter_cost[w] = cost - 0.005 * viol;      /* line_search_filter.hpp:16-17 */
    o->filter_viol[w] = (1 - 0.005) * viol;
    o->filter_n = w + 1;
  }
}
/* computeSolution (line_search.hpp:130-158); alpha = 0: the current point itself */
static void make_trial(stage_t* st, int kind, double alpha) {
  if (alpha == 0.0) {
    memcpy(st->tq, st->q, sizeof(st->tq)); memcpy(st->tv, st->v, sizeof(st->tv)); memcpy(st->ta, st->a, sizeof(st->ta));
    memcpy(st->tu, st->u, sizeof(st->tu)); memcpy(st->tf, st->f, sizeof(st->tf));
    return;
  }
  fb_integrate(st->q, st->dq, alpha, st->tq);
  for (int j = 0; j < NV; ++j) st->tv[j] = fma(alpha, st->dv[j], st->v[j]);
  if (kind == K_TERMINAL) return;
  for (int j = 0; j < NV; ++j) st->ta[j] = fma(alpha, st->daf[j], st->a[j]);
  if (kind != K_IMPULSE)
    for (int j = 0; j < NU; ++j) st->tu[j] = fma(alpha, st->du[j], st->u[j]);
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    for (int x = 0; x < 3; ++x) st->tf[i][x] = st->active[i] ? fma(alpha, st->daf[NV + 3 * k + x], st->f[i][x]) : st->f[i][x];
    if (st->active[i]) ++k;
  }
}
/* SplitOCP::stageCost (split_ocp.hxx:282-298), ImpulseSplitOCP::stageCost, TerminalOCP::terminalCost */
static double trial_stage_cost(const oracle_fb_problem_t* p, const stage_t* st, int kind, double dt, double alpha) {
  const double* wq = kind == K_TERMINAL ? p->qf_weight : (kind == K_IMPULSE ? p->qi_weight : p->q_weight);
  const double* wv = kind == K_TERMINAL ? p->vf_weight : (kind == K_IMPULSE ? p->vi_weight : p->v_weight);
  const double* wa = kind == K_IMPULSE ? p->dvi_weight : p->a_weight;
  const double half = (kind == K_TERMINAL || kind == K_IMPULSE) ? 0.5 : 0.5 * dt;
  double qdiff[NV], l = 0.0;
  fb_subtract(st->tq, st->ref_q, qdiff);
  for (int j = 0; j < NV; ++j) l = fma(wq[j] * qdiff[j], qdiff[j], l);
  for (int j = 0; j < NV; ++j) { const double d = st->tv[j] - st->ref_v[j]; l = fma(wv[j] * d, d, l); }
  if (kind != K_TERMINAL)
    for (int j = 0; j < NV; ++j) l = fma(wa[j] * st->ta[j], st->ta[j], l);
  double cost = half * l;
  if (kind == K_TERMINAL) return cost;
  const double* fw = kind == K_IMPULSE ? p->fi_weight : p->f_weight;
  const double* fr = kind == K_IMPULSE ? p->fi_ref : p->f_ref;
  double lf = 0.0;
  for (int i = 0; i < FB_NC; ++i)
    if (st->active[i])
      for (int x = 0; x < 3; ++x) { const double d = st->tf[i][x] - fr[3 * i + x]; lf = fma(fw[3 * i + x] * d, d, lf); }
  cost += half * lf;
  double bc = 0.0;
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c]) continue;
    double sl = 0.0;
    for (int j = 0; j < comp_rows(p, c); ++j)
      sl += oracle_canon_log(alpha > 0.0 ? fma(alpha, st->c[c].dslack[j], st->c[c].slack[j]) : st->c[c].slack[j]);
    bc += -p->barrier * sl;
  }
  cost += (kind == K_IMPULSE ? 1.0 : dt) * bc;
  return cost;
}
/* SplitOCP::constraintViolation (split_ocp.hxx:301-346), ImpulseSplitOCP::constraintViolation (:178-194) at the
 * trial point; (q_next, v_next) of the trial point of the closing stage */
static double trial_violation(const oracle_fb_problem_t* p, stage_t* st, const elem_t* e, const double* q_next,
                              const double* v_next, const int* imp_active, const double (*ipoints)[3]) {
  const int impulse = e->kind == K_IMPULSE;
  const double dt = e->dt;
  /* constraints: residual with the trial solution and the current slack */
  double cl1 = 0.0;
  for (int c = 0; c < NCOMP; ++c) {
    if (!st->cactive[c] || c == C_DISTANCE) continue;   /* ContactDistance: below, once the trial kinematics exist */
    double s1 = 0.0;
    if (is_cone(c)) {
      const int rpc = cone_rows(p, c);
      for (int i = 0; i < FB_NC; ++i) {
        double r[5];
        if (st->active[i]) friction_residual(p->mu, rpc == 2, st->tf[i], r);
        for (int x = 0; x < rpc; ++x) s1 += st->active[i] ? fabs(r[x] + st->c[c].slack[rpc * i + x]) : 0.0;
      }
    } else {
      for (int j = 0; j < NU; ++j) s1 += fabs(limit_residual_at(p, c, st->tq, st->tv, st->ta, st->tu, j, st->c[c].slack[j]));
    }
    cl1 += s1;
  }
  /* state equation */
  double Fq[NV], fx = 0.0;
  fb_subtract(st->tq, q_next, Fq);
  for (int j = 0; j < NV; ++j) { const double r = impulse ? Fq[j] : fma(dt, st->tv[j], Fq[j]); fx += fabs(r); }
  for (int j = 0; j < NV; ++j) {
    const double r = impulse ? (st->tv[j] + st->ta[j]) - v_next[j] : fma(dt, st->ta[j], st->tv[j]) - v_next[j];
    fx += fabs(r);
  }
  /* (impulse) inverse dynamics and contact rows */
  double f[FB_NC][3], IDC[NVF], idl1 = 0.0;
  fb_kin_t kin;
  for (int i = 0; i < FB_NC; ++i)
    for (int x = 0; x < 3; ++x) f[i][x] = st->active[i] ? st->tf[i][x] : 0.0;
  if (!impulse) {
    fb_forward_kinematics(st->tq, st->tv, st->ta, &kin);
    fb_rnea_derivatives(&kin, f, ANYMAL_GRAVITY, IDC, NULL, NULL, NULL);
    for (int j = 0; j < NU; ++j) IDC[6 + j] -= st->tu[j];
  } else {
    double vpdv[NV];
    fb_forward_kinematics(st->tq, NULL, st->ta, &kin);
    fb_rnea_derivatives(&kin, f, 0.0, IDC, NULL, NULL, NULL);
    for (int j = 0; j < NV; ++j) vpdv[j] = st->tv[j] + st->ta[j];
    fb_forward_kinematics(st->tq, vpdv, NULL, &kin);
  }
  int k = 0;
  for (int i = 0; i < FB_NC; ++i) {
    if (!st->active[i]) continue;
    fb_frame_t fr;
    fb_frame_kinematics(&kin, i, 0, &fr);
    if (!impulse) fb_baumgarte_residual(&fr, p->T / p->N, st->cpoints[i], IDC + NV + 3 * k);
    else for (int x = 0; x < 3; ++x) IDC[NV + 3 * k + x] = fr.vF[x];
    ++k;
  }
  for (int j = 0; j < NV + st->dimf; ++j) idl1 += fabs(IDC[j]);
  if (!impulse && st->cactive[C_DISTANCE]) {   /* computePrimalAndDualResidual at the trial configuration (contact_distance.cpp:134-150) */
    double s1 = 0.0;
    for (int i = 0; i < FB_NC; ++i) {
      if (st->active[i]) continue;
      double P[3];
      fb_contact_point(&kin, i, P);
      s1 += fabs(-P[2] + st->c[C_DISTANCE].slack[i]);
    }
    cl1 += s1;
  }
  if (impulse) return (cl1 + fx) + idl1;
  double viol = (fx + dt * idl1) + dt * cl1;
  if (e->ls_impulse >= 0) {
    /* computeSwitchingConstraintResidual (forward_switching_constraint.hxx:57-68) at the trial point */
    double dqv[NV], q2[NQ], pl1 = 0.0;
    const double c1 = e->dt + e->ls_dt_next, c2 = e->dt * e->ls_dt_next;
    for (int j = 0; j < NV; ++j) dqv[j] = fma(c2, st->ta[j], c1 * st->tv[j]);
    fb_integrate(st->tq, dqv, 1.0, q2);
    fb_forward_kinematics(q2, NULL, NULL, &kin);
    for (int i = 0; i < FB_NC; ++i) {
      if (!imp_active[i]) continue;
      double P[3];
      fb_contact_point(&kin, i, P);
      for (int x = 0; x < 3; ++x) pl1 += fabs(P[x] - ipoints[i][x]);
    }
    viol += pl1;
  }
  return viol;
}
/* computeCostAndViolation + totalCosts / totalViolations: sums in the reference's order (grid stages with the
 * terminal stage, impulse, aux, lift) */
static void ls_cost_and_violation(oracle_fb_ocp_t* o, double alpha, double* cost, double* viol) {
  const int n = o->n_elems;
  for (int e = 0; e < n; ++e) make_trial(&o->slots[o->elems[e].slot], o->elems[e].kind, alpha);
#pragma omp parallel for schedule(dynamic) num_threads(o->nthreads)
  for (int e = 0; e < n; ++e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    st->ls_cost = trial_stage_cost(&o->p, st, el->kind, el->dt, alpha);
    st->ls_viol = 0.0;
    if (el->kind == K_TERMINAL) continue;
    const stage_t* nx = &o->slots[o->elems[el->ls_next].slot];
    int act[HY_MAX_CONTACTS] = {0};
    double pts[HY_MAX_CONTACTS * 3] = {0}, time;
    double ip[FB_NC][3];
    if (el->ls_impulse >= 0) oracle_cs_get_impulse(o->cs, el->ls_impulse, act, pts, &time);
    for (int i = 0; i < FB_NC; ++i)
      for (int x = 0; x < 3; ++x) ip[i][x] = pts[3 * i + x];
    st->ls_viol = trial_violation(&o->p, st, el, nx->tq, nx->tv, act, (const double (*)[3])ip);
  }
  double c = 0.0, v = 0.0;
  for (int pass = 0; pass < 4; ++pass) {
    double cs = 0.0, vs = 0.0;
    for (int e = 0; e < n; ++e) {
      const int k = o->elems[e].kind;
      const int mine = (pass == 0 && (k == K_GRID || k == K_TERMINAL)) || (pass == 1 && k == K_IMPULSE) ||
                       (pass == 2 && k == K_AUX) || (pass == 3 && k == K_LIFT);
      if (!mine) continue;
      cs += o->slots[o->elems[e].slot].ls_cost;
      vs += o->slots[o->elems[e].slot].ls_viol;
    }
    c = pass == 0 ? cs : c + cs;
    v = pass == 0 ? vs : v + vs;
  }
  *cost = c;
  *viol = v;
}
static double line_search_step(oracle_fb_ocp_t* o, double max_primal) {
  double cost, viol;
  if (o->filter_n == 0) {
    ls_cost_and_violation(o, 0.0, &cost, &viol);
    filter_augment(o, cost, viol);
  }
  double alpha = max_primal;
  while (alpha > 0.05) {
    ls_cost_and_violation(o, alpha, &cost, &viol);
    if (filter_is_accepted(o, cost, viol)) {
      filter_augment(o, cost, viol);
      break;
    }
    alpha *= 0.75;
  }
  return alpha > 0.05 ? alpha : 0.05;
}
void oracle_fb_ocp_clear_line_search_filter(oracle_fb_ocp_t* o) { o->filter_n = 0; }
int oracle_fb_ocp_filter_size(const oracle_fb_ocp_t* o) { return o->filter_n; }
/* cost / violation totals of the point s + alpha d of the last direction (tests) */
void oracle_fb_ocp_cost_and_violation(oracle_fb_ocp_t* o, double alpha, double* out) { ls_cost_and_violation(o, alpha, &out[0], &out[1]); }

/* OCPSolver::updateSolution (ocp_solver.cpp:67-92) */
int oracle_fb_ocp_update_solution_ls(oracle_fb_ocp_t* o, double t, const double* q, const double* v, int line_search) {
  const int rc = discretize(o, t);
  if (rc == -1 || rc == -2) return rc;
  const int n = o->n_elems;
  linearize_all(o, q, 0);
  /* RiccatiRecursionSolver::backwardRiccatiRecursion (riccati_recursion_solver.cpp:48-107) */
  {
    stage_t* sN = &o->slots[o->elems[n - 1].slot];
    for (int r = 0; r < NV; ++r)
      for (int c = 0; c < NV; ++c) {
        sN->Pqq[r * NV + c] = sN->Qxx[r * NX + c];
        sN->Pvv[r * NV + c] = sN->Qxx[(NV + r) * NX + NV + c];
        sN->Pqv[r * NV + c] = 0.0;
      }
    for (int j = 0; j < NV; ++j) { sN->sq[j] = -sN->lq[j]; sN->sv[j] = -sN->lv[j]; }
  }
  for (int e = n - 2; e >= 0; --e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    ric_t rn;
    ric_of(&o->slots[o->elems[e + 1].slot], &rn);
    riccati_backward_stage(st, el->kind == K_IMPULSE, el->dt, &rn, el->sw_impulse >= 0);
  }
  /* computeInitialStateDirection (:110-126) */
  {
    stage_t* s0 = &o->slots[o->elems[0].slot];
    double d6[6];
    fb_subtract(q, s0->q, s0->dq);
    mv(MM_SET, 6, 6, s0->Fqq_prev_inv, 6, 1, s0->dq, d6);
    for (int j = 0; j < 6; ++j) s0->dq[j] = -d6[j];
    for (int j = 0; j < NV; ++j) s0->dv[j] = v[j] - s0->v[j];
  }
  /* forwardRiccatiRecursion (:129-162) */
  for (int e = 0; e + 1 < n; ++e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    stage_t* sn = &o->slots[o->elems[e + 1].slot];
    riccati_forward_stage(st, el->kind == K_IMPULSE, el->dt, sn->dq, sn->dv);
  }
  /* computeDirection (:165-241) */
#pragma omp parallel for schedule(dynamic) num_threads(o->nthreads)
  for (int e = 0; e < n; ++e) {
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    costate_direction(st);
    if (el->kind == K_TERMINAL) { st->max_primal = 1.0; st->max_dual = 1.0; continue; }
    condensed_primal_direction(st, el->kind == K_IMPULSE);
    slack_dual_direction(&o->p, st);
    if (el->sw_impulse >= 0) {
      double dx[NX];
      memcpy(dx, st->dq, sizeof(st->dq));
      memcpy(dx + NV, st->dv, sizeof(st->dv));
      mv(MM_SET, st->dimi, NX, st->cM, NX, 1, dx, st->dxi);
      for (int j = 0; j < st->dimi; ++j) st->dxi[j] += st->cm[j];
    }
    max_step_sizes(&o->p, st);
  }
  /* min over the reference's stage order: grid 0..N, impulses, aux, lifts (exact, order-free) */
  double ap = 1.0, ad = 1.0;
  for (int e = 0; e < n; ++e) {
    const stage_t* st = &o->slots[o->elems[e].slot];
    if (st->max_primal < ap) ap = st->max_primal;
    if (st->max_dual < ad) ad = st->max_dual;
  }
  if (line_search) ap = line_search_step(o, ap);
  o->primal_step = ap;
  o->dual_step = ad;
  /* OCPLinearizer::integrateSolution (ocp_linearizer.cpp:140-221) */
  for (int e = 0; e < n; ++e) {       /* dual directions need the untouched dgmm of the next stage: two passes */
    const elem_t* el = &o->elems[e];
    stage_t* st = &o->slots[el->slot];
    if (el->kind != K_TERMINAL)
      condensed_dual_direction(st, el->kind == K_IMPULSE, el->dt, o->slots[o->elems[e + 1].slot].dgmm);
    correct_costate_direction(st);
  }
  for (int e = 0; e < n; ++e) update_stage(&o->slots[o->elems[e].slot], o->elems[e].kind, ap, ad);
  int info = 0;
  for (int e = 0; e < n; ++e)
    if (o->slots[o->elems[e].slot].chol_info && !info) info = o->slots[o->elems[e].slot].chol_info;
  return info ? 1000 + info : rc;
}

int oracle_fb_ocp_update_solution(oracle_fb_ocp_t* o, double t, const double* q, const double* v) {
  return oracle_fb_ocp_update_solution_ls(o, t, q, v, 0);
}

/* full-host-core mode of the CPU baseline (BASELINE.md mode B): OpenMP over independent instances, every instance
 * single-threaded (nested parallelism is off, so the stage loops inside run serially) */
void oracle_fb_ocp_batch_update_solution(oracle_fb_ocp_t** os, int batch, double t, const double* q, const double* v,
                                         int line_search, int nthreads) {
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
  for (int b = 0; b < batch; ++b) oracle_fb_ocp_update_solution_ls(os[b], t, q + (size_t)b * NQ, v + (size_t)b * NV, line_search);
}

/* OCPSolver::computeKKTResidual / KKTError (ocp_solver.cpp:202-213, ocp_linearizer.cpp:97-137) */
int oracle_fb_ocp_compute_kkt_residual(oracle_fb_ocp_t* o, double t, const double* q, const double* v) {
  (void)v;
  const int rc = discretize(o, t);
  if (rc == -1 || rc == -2) return rc;
  linearize_all(o, q, 1);
  return rc;
}
double oracle_fb_ocp_kkt_error(const oracle_fb_ocp_t* o) {
  /* sum in the reference's order: grid stages 0..N, impulse, aux, lift */
  double sum = 0.0;
  for (int pass = 0; pass < 4; ++pass)
    for (int e = 0; e < o->n_elems; ++e) {
      const int k = o->elems[e].kind;
      const int mine = (pass == 0 && (k == K_GRID || k == K_TERMINAL)) || (pass == 1 && k == K_IMPULSE) ||
                       (pass == 2 && k == K_AUX) || (pass == 3 && k == K_LIFT);
      if (mine) sum += o->slots[o->elems[e].slot].kkt_sq;
    }
  return sqrt(sum);
}
void oracle_fb_ocp_get_step_sizes(const oracle_fb_ocp_t* o, double* out) { out[0] = o->primal_step; out[1] = o->dual_step; }

/* getters by chain position e.  name: q v a u f(12, per contact) lmd gmm beta mu(12) nu_passive xi(12)
 * and directions dq dv du daf(30) dbetamu(30) dlmd dgmm dnu_passive dxi; "kkt" the stage's squared KKT norm */
int oracle_fb_ocp_get(const oracle_fb_ocp_t* o, int e, const char* name, double* out) {
  if (e < 0 || e >= o->n_elems) return -1;
  const stage_t* st = &o->slots[o->elems[e].slot];
#define GET(nm, field) if (!strcmp(name, nm)) { memcpy(out, st->field, sizeof(st->field)); return (int)(sizeof(st->field) / sizeof(double)); }
  GET("q", q) GET("v", v) GET("a", a) GET("u", u) GET("f", f) GET("lmd", lmd) GET("gmm", gmm) GET("beta", beta) GET("mu", mu)
  GET("nu_passive", nu_passive) GET("xi", xi) GET("dq", dq) GET("dv", dv) GET("du", du) GET("daf", daf) GET("dbetamu", dbetamu)
  GET("dlmd", dlmd) GET("dgmm", dgmm) GET("dnu_passive", dnu_passive) GET("dxi", dxi)
  GET("lq", lq) GET("lv", lv) GET("la", la) GET("lf", lf) GET("lu", lu) GET("lu_passive", lu_passive) GET("Fq", Fq) GET("Fv", Fv)
  GET("P", P) GET("IDC", IDC) GET("Qxx", Qxx) GET("Qxu", Qxu) GET("Quu", Quu) GET("Qaa", Qaa) GET("Qff", Qff) GET("Fvq", Fvq) GET("Fvv", Fvv)
  GET("Fvu", Fvu) GET("Fqq6", Fqq6) GET("Fqv6", Fqv6) GET("MJtJinv", MJtJinv) GET("MJ_dIDC", MJ_dIDC) GET("MJ_IDC", MJ_IDC)
  GET("dIDCdqv", dIDCdqv) GET("dCda", dCda) GET("Mm", Mm) GET("K", K) GET("k", k) GET("Pqq", Pqq) GET("Pqv", Pqv) GET("Pvv", Pvv)
  GET("sq", sq) GET("sv", sv) GET("Phix", Phix) GET("Phia", Phia) GET("Phiu", Phiu) GET("cM", cM) GET("cm", cm)
  GET("Fqq_prev_inv", Fqq_prev_inv) GET("Fqq_inv", Fqq_inv) GET("laf", laf) GET("Qafqv", Qafqv) GET("Qafu", Qafu)
#undef GET
  if (!strcmp(name, "kkt")) { out[0] = st->kkt_sq; return 1; }
  if (!strcmp(name, "cdJ")) { memcpy(out, st->cdJ, sizeof(st->cdJ)); return FB_NC * NV; }
  if (!strcmp(name, "cdz")) { memcpy(out, st->cdz, sizeof(st->cdz)); return FB_NC; }
  if (!strcmp(name, "active")) { for (int i = 0; i < FB_NC; ++i) out[i] = st->active[i]; return FB_NC; }
  if (!strcmp(name, "ls_cost")) { out[0] = st->ls_cost; return 1; }
  if (!strcmp(name, "ls_viol")) { out[0] = st->ls_viol; return 1; }
  if (!strncmp(name, "slack", 5) || !strncmp(name, "dual", 4)) {
    int n = 0;
    for (int c = 0; c < NCOMP; ++c) {
      const double* src = name[0] == 's' ? st->c[c].slack : st->c[c].dual;
      for (int j = 0; j < comp_dim(c); ++j) out[n++] = st->cactive[c] ? src[j] : 0.0;
    }
    return n;
  }
  return -1;
}
int oracle_fb_ocp_set(oracle_fb_ocp_t* o, int e, const char* name, const double* in) {
  if (e < 0 || e >= o->n_elems) return -1;
  stage_t* st = &o->slots[o->elems[e].slot];
#define SET(nm, field) if (!strcmp(name, nm)) { memcpy(st->field, in, sizeof(st->field)); return 0; }
  SET("q", q) SET("v", v) SET("a", a) SET("u", u) SET("f", f) SET("lmd", lmd) SET("gmm", gmm) SET("beta", beta) SET("mu", mu)
  SET("nu_passive", nu_passive) SET("xi", xi)
#undef SET
  return -1;
}

/* batch helpers of the test harness (OpenMP over instances): KKT errors and one field of one chain element of
 * every instance, out[batch][size] */
void oracle_fb_ocp_batch_kkt(oracle_fb_ocp_t** os, int batch, double t, const double* q, const double* v, double* kkt_out,
                             int nthreads) {
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
  for (int b = 0; b < batch; ++b) {
    oracle_fb_ocp_compute_kkt_residual(os[b], t, q + (size_t)b * NQ, v + (size_t)b * NV);
    kkt_out[b] = oracle_fb_ocp_kkt_error(os[b]);
  }
}
int oracle_fb_ocp_batch_get(oracle_fb_ocp_t** os, int batch, int e, const char* name, double* out, int size) {
  int n = 0;
  for (int b = 0; b < batch; ++b) {
    n = oracle_fb_ocp_get(os[b], e, name, out + (size_t)b * size);
    if (n != size) return n;
  }
  return n;
}


/* oracle/fb_oracle.c -- TEST INFRASTRUCTURE: CPU oracle of the floating-base (ANYmal) OCPSolver path,
 * SURVEY.md §8 row a12.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * "Parity unpinned": pinocchio/Eigen are absent, the reference cannot be built and ships no golden vectors;
 * the restatement is validated by finite differences, an independent body-frame RNEA and the identities of the
 * reference's own unit tests (tests/test_oracle_fb_*.py). */
#include <stdlib.h>
#include <stdio.h>

#include "fb_robot.h"

/* ---------------------------------------------------------------------------------------------- */
/* exports of the robot layer for the tests                                                        */
/* ---------------------------------------------------------------------------------------------- */
void oracle_fb_integrate(const double* q, const double* v, double alpha, double* out) { fb_integrate(q, v, alpha, out); }
void oracle_fb_subtract(const double* qp, const double* qm, double* out) { fb_subtract(qp, qm, out); }
void oracle_fb_dsubtract(const double* qp, const double* qm, double* Jplus, double* Jminus) {
  fb_dsubtract_dplus(qp, qm, Jplus);
  fb_dsubtract_dminus(qp, qm, Jminus);
}
void oracle_fb_dsubtract_inverse(const double* J, double* Jinv) { fb_dsubtract_inverse(J, Jinv); }
void oracle_fb_dintegrate(const double* v6, double* Jq, double* Jv) {
  fb_dintegrate_dq(v6, Jq);
  fb_dintegrate_dv(v6, Jv);
}
void oracle_fb_exp6(const double* nu, double* R, double* p) { fb_exp6(nu, R, p); }
void oracle_fb_log6(const double* R, const double* p, double* out) { fb_log6(R, p, out); }
void oracle_fb_rnea(const double* q, const double* v, const double* a, const double* f12, double gravity,
                    double* tau, double* dq, double* dv, double* M) {
  fb_kin_t k;
  fb_forward_kinematics(q, v, a, &k);
  double f[FB_NC][3];
  for (int i = 0; i < FB_NC; ++i)
    for (int e = 0; e < 3; ++e) f[i][e] = f12 ? f12[3 * i + e] : 0.0;
  fb_rnea_derivatives(&k, f, gravity, tau, dq, dv, M);
}
/* frame outputs: P[3], vF[6], aF[6], J[6x18], v_dq, a_dq, a_dv; Baumgarte: C[3], dCdq/dv/da [3x18] */
void oracle_fb_contact(const double* q, const double* v, const double* a, int contact, double time_step,
                       const double* contact_point, double* P, double* vF, double* aF, double* J, double* v_dq,
                       double* a_dq, double* a_dv, double* C, double* dCdq, double* dCdv, double* dCda) {
  fb_kin_t k;
  fb_frame_t fr;
  fb_forward_kinematics(q, v, a, &k);
  fb_frame_kinematics(&k, contact, 2, &fr);
  memcpy(P, fr.P, sizeof(fr.P));
  memcpy(vF, fr.vF, sizeof(fr.vF));
  memcpy(aF, fr.aF, sizeof(fr.aF));
  memcpy(J, fr.J, sizeof(fr.J));
  memcpy(v_dq, fr.v_dq, sizeof(fr.v_dq));
  memcpy(a_dq, fr.a_dq, sizeof(fr.a_dq));
  memcpy(a_dv, fr.a_dv, sizeof(fr.a_dv));
  fb_baumgarte_residual(&fr, time_step, contact_point, C);
  fb_baumgarte_derivatives(&k, contact, &fr, time_step, dCdq, dCdv, dCda);
}
int oracle_fb_mjtjinv(const double* M, const double* J, int dimf, double* out) {
  return fb_MJtJinv(M, J, dimf, out, FB_NV + dimf);
}

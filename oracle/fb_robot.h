/* oracle/fb_robot.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product.
 *
 * Floating-base rigid-body layer of the oracle for the ANYmal path: everything the
 * reference's Robot class obtains from pinocchio on the OCPSolver hot path, restated from the published
 * algorithms (pinocchio itself is not in /root/reference: "parity unpinned"):
 *
 *   Robot::updateKinematics / RNEA / RNEADerivatives / RNEAImpulse / RNEAImpulseDerivatives
 *                                                    include/idocp/robot/robot.hxx:193-230,444-535
 *   Robot::setContactForces / setImpulseForces       robot.hxx:408-438, point_contact.hxx:15-20
 *   Robot::computeBaumgarteResidual/Derivatives, computeImpulseVelocityResidual/Derivatives,
 *   computeContactResidual/Derivative                robot.hxx:246-350, point_contact.hxx:67-205
 *   Robot::computeMJtJinv                            robot.hxx:576-615
 *   Robot::integrateConfiguration, subtractConfiguration, dSubtractdConfiguration{Plus,Minus,Inverse},
 *   dIntegratedConfiguration/Velocity                robot.hxx:22-170
 *
 * Conventions (pinocchio's): q = [p(3), quaternion xyzw, 12 joint angles], v = [base linear, base angular
 * (both in the base frame), 12 joint rates]; spatial vectors are [linear; angular]; all recursions are
 * written in the WORLD frame (the formulation of pinocchio::computeRNEADerivatives).  Matrices are
 * row-major.  CANONICAL ARITHMETIC: compiled with -ffp-contract=off, every fused multiply-add is an
 * explicit fma(); the CUDA kernels (idocp_b200/csrc/fb_*.cuh) execute the same operation trees. */
#ifndef ORACLE_FB_ROBOT_H_
#define ORACLE_FB_ROBOT_H_

#include <math.h>
#include <string.h>

#include "model_anymal.h"
#include "canon_pivot.h"

#define FB_NV 18
#define FB_NQ 19
#define FB_NU 12
#define FB_NB 13
#define FB_NC 4
#define FB_MAXF 12
#define FB_TAYLOR 1.220703125e-04 /* TaylorSeriesExpansion<double>::precision<3>() = 2^-13 */
#define FB_PI 3.14159265358979311600e+00

/* shared canonical elementary functions (oracle/idocp_oracle.c) */
void oracle_canon_sincos(double x, double* sn, double* cs);
double oracle_canon_acos(double x);
double oracle_canon_log(double x);

/* ---------------------------------------------------------------------------------------------- */
/* small vectors                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
static inline void fb_cross(const double* a, const double* b, double* c) { /* c must not alias a, b */
  c[0] = fma(a[1], b[2], -(a[2] * b[1]));
  c[1] = fma(a[2], b[0], -(a[0] * b[2]));
  c[2] = fma(a[0], b[1], -(a[1] * b[0]));
}
static inline double fb_dot3(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
/* y = R x  /  y = R^T x, R row-major 3x3, y must not alias x */
static inline void fb_rot(const double* R, const double* x, double* y) {
  for (int i = 0; i < 3; ++i) y[i] = fma(R[3 * i + 2], x[2], fma(R[3 * i + 1], x[1], R[3 * i] * x[0]));
}
static inline void fb_rotT(const double* R, const double* x, double* y) {
  for (int i = 0; i < 3; ++i) y[i] = fma(R[6 + i], x[2], fma(R[3 + i], x[1], R[i] * x[0]));
}
/* C = A B, 3x3 row-major */
static inline void fb_mul33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(A[3 * r + 2], B[6 + k], fma(A[3 * r + 1], B[3 + k], A[3 * r] * B[k]));
}
/* C = A^T B */
static inline void fb_mulT33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(A[6 + r], B[6 + k], fma(A[3 + r], B[3 + k], A[r] * B[k]));
}
/* M += skew(v) */
static inline void fb_add_skew(const double* v, double* M) {
  M[1] -= v[2]; M[2] += v[1]; M[3] += v[2]; M[5] -= v[0]; M[6] -= v[1]; M[7] += v[0];
}
/* y = A x, A symmetric stored (xx,xy,xz,yy,yz,zz) */
static inline void fb_sym3(const double* A, const double* x, double* y) {
  y[0] = fma(A[2], x[2], fma(A[1], x[1], A[0] * x[0]));
  y[1] = fma(A[4], x[2], fma(A[3], x[1], A[1] * x[0]));
  y[2] = fma(A[5], x[2], fma(A[4], x[1], A[2] * x[0]));
}

/* spatial motion cross product c = a x b; force cross c = a x* f; pairing <m, f> */
static inline void fb_mxm(const double* a, const double* b, double* c) {
  double t1[3], t2[3];
  fb_cross(a + 3, b, t1);
  fb_cross(a, b + 3, t2);
  for (int i = 0; i < 3; ++i) c[i] = t1[i] + t2[i];
  fb_cross(a + 3, b + 3, c + 3);
}
static inline void fb_mxf(const double* a, const double* f, double* c) {
  double t1[3], t2[3];
  fb_cross(a + 3, f, c);
  fb_cross(a + 3, f + 3, t1);
  fb_cross(a, f, t2);
  for (int i = 0; i < 3; ++i) c[3 + i] = t1[i] + t2[i];
}
static inline double fb_dot6(const double* m, const double* f) {
  double acc = m[0] * f[0];
  for (int i = 1; i < 6; ++i) acc = fma(m[i], f[i], acc);
  return acc;
}

/* spatial inertia about the world origin: mass, first moment h = m c, rotational inertia about the origin */
typedef struct { double m, h[3], I[6]; } fb_inertia_t;
/* f = Y mv : lin = m u - h x w, ang = I w + h x u */
static inline void fb_Ymul(const fb_inertia_t* Y, const double* mv, double* f) {
  double hw[3], hu[3], Iw[3];
  fb_cross(Y->h, mv + 3, hw);
  fb_cross(Y->h, mv, hu);
  fb_sym3(Y->I, mv + 3, Iw);
  for (int i = 0; i < 3; ++i) {
    f[i] = fma(Y->m, mv[i], -hw[i]);
    f[3 + i] = Iw[i] + hu[i];
  }
}
/* "doYcrb" of pinocchio::computeRNEADerivatives, D m = v x* (Y m) - Y (v x m) + m x* (Y v).  Its first
 * three columns vanish: D = [[0, -2 [pl]x], [0, Sym - [pa]x]] with (pl, pa) = Y v the momentum. */
typedef struct { double pl[3], pa[3], S[6]; } fb_dinertia_t;
static inline void fb_dinertia(const fb_inertia_t* Y, const double* v, fb_dinertia_t* D) {
  double mom[6];
  fb_Ymul(Y, v, mom);
  for (int i = 0; i < 3; ++i) { D->pl[i] = mom[i]; D->pa[i] = mom[3 + i]; }
  /* A = [va]x I (columns va x I_col), Sym = A + A^T - (h vl^T + vl h^T) + 2 (vl.h) Id */
  const double* va = v + 3;
  const double* vl = v;
  const double I0[3] = {Y->I[0], Y->I[1], Y->I[2]}, I1[3] = {Y->I[1], Y->I[3], Y->I[4]}, I2[3] = {Y->I[2], Y->I[4], Y->I[5]};
  double A[3][3], c[3];
  fb_cross(va, I0, c); A[0][0] = c[0]; A[1][0] = c[1]; A[2][0] = c[2];
  fb_cross(va, I1, c); A[0][1] = c[0]; A[1][1] = c[1]; A[2][1] = c[2];
  fb_cross(va, I2, c); A[0][2] = c[0]; A[1][2] = c[1]; A[2][2] = c[2];
  const double d2 = 2.0 * fb_dot3(vl, Y->h);
  const int ii[6] = {0, 0, 0, 1, 1, 2}, jj[6] = {0, 1, 2, 1, 2, 2};
  for (int k = 0; k < 6; ++k) {
    const int i = ii[k], j = jj[k];
    double s = (A[i][j] + A[j][i]) - fma(Y->h[i], vl[j], vl[i] * Y->h[j]);
    if (i == j) s += d2;
    D->S[k] = s;
  }
}
/* y = D m  and  y = D^T m */
static inline void fb_Dmul(const fb_dinertia_t* D, const double* m, double* y) {
  double t[3], s[3], u[3];
  fb_cross(m + 3, D->pl, t);          /* w x pl */
  fb_sym3(D->S, m + 3, s);
  fb_cross(D->pa, m + 3, u);          /* pa x w */
  for (int i = 0; i < 3; ++i) { y[i] = 2.0 * t[i]; y[3 + i] = s[i] - u[i]; }
}
static inline void fb_DTmul(const fb_dinertia_t* D, const double* m, double* y) {
  double t[3], s[3], u[3];
  fb_cross(D->pl, m, t);              /* pl x u */
  fb_sym3(D->S, m + 3, s);
  fb_cross(D->pa, m + 3, u);          /* pa x w */
  for (int i = 0; i < 3; ++i) { y[i] = 0.0; y[3 + i] = fma(2.0, t[i], s[i]) + u[i]; }
}

/* ---------------------------------------------------------------------------------------------- */
/* tree structure                                                                                  */
/* ---------------------------------------------------------------------------------------------- */
static inline int fb_body_of_dof(int c) { return c < 6 ? 0 : c - 5; }
static inline int fb_parent_body(int b) { return ANYMAL_JOINT_PARENT[b - 1] + 1; } /* b >= 1 */
/* 1 when dof r belongs to a strict ancestor joint of dof c */
static inline int fb_is_ancestor(int r, int c) {
  if (r < 6) return c >= 6;
  if (c < 6) return 0;
  return ((r - 6) / 3 == (c - 6) / 3) && r < c;
}
static inline int fb_same_joint(int r, int c) { return (r < 6 && c < 6) || r == c; }

/* ---------------------------------------------------------------------------------------------- */
/* SO(3) / SE(3): pinocchio::exp3/log3/Jlog3/exp6/log6/Jlog6/Jexp6 (explog.hpp), Eigen quaternions      */
/* ---------------------------------------------------------------------------------------------- */
/* Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w) */
static inline void fb_quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
/* Eigen: Quaternion = rotation matrix (Shoemake) */
static inline void fb_R_to_quat(const double* R, double* q) {
  double t = (R[0] + R[4]) + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(((R[4 * i] - R[4 * j]) - R[4 * k]) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

static inline void fb_log3(const double* R, double* theta_out, double* w) {
  const double tr = (R[0] + R[4]) + R[8];
  double theta;
  if (tr > 3.0) theta = 0.0;
  else if (tr < -1.0) theta = FB_PI;
  else theta = oracle_canon_acos((tr - 1.0) * 0.5);
  *theta_out = theta;
  if (theta >= FB_PI - 1e-2) {
    double sn, cphi;
    oracle_canon_sincos(theta - FB_PI, &sn, &cphi);
    const double beta = (theta * theta) / (1.0 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0.0 ? sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0.0 ? sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0.0 ? sqrt(t2) : 0.0);
    return;
  }
  double t = 1.0;
  if (theta > FB_TAYLOR) {
    double sn, cs;
    oracle_canon_sincos(theta, &sn, &cs);
    t = theta / sn;
  }
  t *= 0.5;
  w[0] = t * (R[7] - R[5]); w[1] = t * (R[2] - R[6]); w[2] = t * (R[3] - R[1]);
}
static inline void fb_Jlog3(double theta, const double* w, double* A) {
  double alpha, diag;
  if (theta < FB_TAYLOR) {
    alpha = 1.0 / 12.0 + (theta * theta) / 720.0;
    diag = 0.5 * (2.0 - (theta * theta) / 6.0);
  } else {
    double st, ct;
    oracle_canon_sincos(theta, &st, &ct);
    const double st_1mct = st / (1.0 - ct);
    alpha = 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
    diag = 0.5 * (theta * st_1mct);
  }
  double aw[3] = {alpha * w[0], alpha * w[1], alpha * w[2]}, hw[3] = {0.5 * w[0], 0.5 * w[1], 0.5 * w[2]};
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) A[3 * r + k] = aw[r] * w[k];
  A[0] += diag; A[4] += diag; A[8] += diag;
  fb_add_skew(hw, A);
}
/* log6 of the placement (R, p): out = [v; w] */
static inline void fb_log6(const double* R, const double* p, double* out) {
  double theta, w[3];
  fb_log3(R, &theta, w);
  const double t2 = theta * theta;
  double alpha, beta;
  if (theta < FB_TAYLOR) {
    alpha = (1.0 - t2 / 12.0) - (t2 * t2) / 720.0;
    beta = 1.0 / 12.0 + t2 / 720.0;
  } else {
    double st, ct;
    oracle_canon_sincos(theta, &st, &ct);
    alpha = (theta * st) / (2.0 * (1.0 - ct));
    beta = 1.0 / t2 - st / ((2.0 * theta) * (1.0 - ct));
  }
  double wxp[3];
  fb_cross(w, p, wxp);
  const double bwp = beta * fb_dot3(w, p);
  for (int i = 0; i < 3; ++i) {
    out[i] = fma(bwp, w[i], fma(-0.5, wxp[i], alpha * p[i]));
    out[3 + i] = w[i];
  }
}
/* the (beta, beta_dot_over_theta) pair and the C block shared by Jlog6 and Jexp6 */
static inline void fb_se3_C(double theta, const double* w, const double* p, double* C) {
  const double t2 = theta * theta;
  double beta, bdot;
  if (theta < FB_TAYLOR) {
    beta = 1.0 / 12.0 + t2 / 720.0;
    bdot = 1.0 / 360.0;
  } else {
    double st, ct;
    oracle_canon_sincos(theta, &st, &ct);
    const double tinv = 1.0 / theta, t2inv = tinv * tinv;
    const double inv_2_2ct = 1.0 / (2.0 * (1.0 - ct));
    beta = t2inv - (st * tinv) * inv_2_2ct;
    bdot = -2.0 * (t2inv * t2inv) + ((1.0 + st * tinv) * t2inv) * inv_2_2ct;
  }
  const double wTp = fb_dot3(w, p);
  const double c1 = bdot * wTp, c2 = fma(t2, bdot, 2.0 * beta);
  double v3[3], bw[3], hp[3];
  for (int i = 0; i < 3; ++i) { v3[i] = c1 * w[i] - c2 * p[i]; bw[i] = beta * w[i]; hp[i] = 0.5 * p[i]; }
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(bw[r], p[k], v3[r] * w[k]);
  const double dg = wTp * beta;
  C[0] += dg; C[4] += dg; C[8] += dg;
  fb_add_skew(hp, C);
}
/* Jlog6 of the placement (R, p): J = [[A, B], [0, A]], row-major 6x6 */
static inline void fb_Jlog6(const double* R, const double* p, double* J) {
  double theta, w[3], A[9], B[9], C[9];
  fb_log3(R, &theta, w);
  fb_Jlog3(theta, w, A);
  fb_se3_C(theta, w, p, C);
  fb_mul33(C, A, B);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      J[6 * r + k] = A[3 * r + k];
      J[6 * r + 3 + k] = B[3 * r + k];
      J[6 * (3 + r) + k] = 0.0;
      J[6 * (3 + r) + 3 + k] = A[3 * r + k];
    }
}
/* exp6([v; w]) -> (R, p) */
static inline void fb_exp6(const double* nu, double* R, double* p) {
  const double* v = nu;
  const double* w = nu + 3;
  const double t2 = fb_dot3(w, w);
  const double t = sqrt(t2);
  double alpha_wxv, alpha_v, alpha_w, diag;
  if (t > FB_TAYLOR) {
    double st, ct;
    oracle_canon_sincos(t, &st, &ct);
    const double inv_t2 = 1.0 / t2;
    alpha_wxv = (1.0 - ct) * inv_t2;
    alpha_v = st / t;
    alpha_w = ((1.0 - alpha_v) * inv_t2) * fb_dot3(w, v);
    diag = ct;
  } else {
    alpha_wxv = 0.5 - t2 / 24.0;
    alpha_v = 1.0 - t2 / 6.0;
    alpha_w = (1.0 / 6.0 - t2 / 120.0) * fb_dot3(w, v);
    diag = 1.0 - t2 / 2.0;
  }
  double wxv[3];
  fb_cross(w, v, wxv);
  for (int i = 0; i < 3; ++i) p[i] = fma(alpha_wxv, wxv[i], fma(alpha_w, w[i], alpha_v * v[i]));
  double aw[3] = {alpha_wxv * w[0], alpha_wxv * w[1], alpha_wxv * w[2]}, avw[3] = {alpha_v * w[0], alpha_v * w[1], alpha_v * w[2]};
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) R[3 * r + k] = aw[r] * w[k];
  fb_add_skew(avw, R);
  R[0] += diag; R[4] += diag; R[8] += diag;
}
static inline void fb_Jexp3(const double* r, double* J) {
  const double n2 = fb_dot3(r, r);
  const double n = sqrt(n2);
  double a, b, c;
  if (n < FB_TAYLOR) {
    a = 1.0 - n2 / 6.0;
    b = -0.5 - n2 / 24.0;
    c = 1.0 / 6.0 - n2 / 120.0;
  } else {
    double sn, cn;
    oracle_canon_sincos(n, &sn, &cn);
    const double n_inv = 1.0 / n, n2_inv = n_inv * n_inv;
    a = sn * n_inv;
    b = -(1.0 - cn) * n2_inv;
    c = n2_inv * (1.0 - a);
  }
  double cr[3] = {c * r[0], c * r[1], c * r[2]}, br[3] = {b * r[0], b * r[1], b * r[2]};
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) J[3 * i + k] = cr[i] * r[k];
  J[0] += a; J[4] += a; J[8] += a;
  fb_add_skew(br, J);   /* J(0,1) = -b r2, J(0,2) = b r1, J(1,2) = -b r0 and the antisymmetric partners */
}
/* Jexp6([v; w]) = [[A, B], [0, A]] (right Jacobian of exp6), row-major 6x6 */
static inline void fb_Jexp6(const double* nu, double* J) {
  const double* v = nu;
  const double* w = nu + 3;
  double A[9], B[9], C[9], p[3];
  fb_Jexp3(w, A);
  fb_rotT(A, v, p);                       /* p = A^T v */
  const double t = sqrt(fb_dot3(w, w));
  /* Jexp6 = Jlog6(exp6(nu))^-1.  With Jlog6 = [[Al, Cl Al], [0, Al]] and Al = A^-1 the inverse is
   * [[A, -A Cl], [0, A]], Cl being the C block at the translation of exp6(nu), which is p = A^T v (the left
   * Jacobian of SO(3) applied to v).  Checked against finite differences in tests/test_oracle_fb_robot.py. */
  fb_se3_C(t, w, p, C);
  fb_mul33(A, C, B);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      J[6 * r + k] = A[3 * r + k];
      J[6 * r + 3 + k] = -B[3 * r + k];
      J[6 * (3 + r) + k] = 0.0;
      J[6 * (3 + r) + 3 + k] = A[3 * r + k];
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* configuration space of the free-flyer + 12 revolute joints (pinocchio joint-configuration.hpp)   */
/* ---------------------------------------------------------------------------------------------- */
/* Robot::integrateConfiguration: q_out = q (+) alpha v  (robot.hxx:22-60).  SE(3): M_out = M exp6(alpha v_base),
 * quaternion from the rotation matrix, sign-aligned with the input quaternion, first-order normalised. */
static inline void fb_integrate(const double* q, const double* v, double alpha, double* q_out) {
  double nu[6], R0[9], Re[9], pe[3], R1[9], quat[4];
  for (int i = 0; i < 6; ++i) nu[i] = alpha * v[i];
  fb_quat_to_R(q + 3, R0);
  fb_exp6(nu, Re, pe);
  fb_mul33(R0, Re, R1);
  for (int i = 0; i < 3; ++i)
    q_out[i] = fma(R0[3 * i + 2], pe[2], fma(R0[3 * i + 1], pe[1], fma(R0[3 * i], pe[0], q[i])));
  fb_R_to_quat(R1, quat);
  double dot = quat[0] * q[3];
  for (int i = 1; i < 4; ++i) dot = fma(quat[i], q[3 + i], dot);
  if (dot < 0.0)
    for (int i = 0; i < 4; ++i) quat[i] = -quat[i];
  double n2 = quat[0] * quat[0];
  for (int i = 1; i < 4; ++i) n2 = fma(quat[i], quat[i], n2);
  const double corr = (3.0 - n2) / 2.0;   /* quaternion::firstOrderNormalize */
  for (int i = 0; i < 4; ++i) q_out[3 + i] = quat[i] * corr;
  for (int j = 0; j < FB_NU; ++j) q_out[7 + j] = fma(alpha, v[6 + j], q[7 + j]);
}
/* relative placement M = M0^-1 M1 of the bases of two configurations */
static inline void fb_relative(const double* q0, const double* q1, double* R, double* p) {
  double R0[9], R1[9], dp[3];
  fb_quat_to_R(q0 + 3, R0);
  fb_quat_to_R(q1 + 3, R1);
  fb_mulT33(R0, R1, R);
  for (int i = 0; i < 3; ++i) dp[i] = q1[i] - q0[i];
  fb_rotT(R0, dp, p);
}
/* Robot::subtractConfiguration(q_plus, q_minus, out): out = q_plus (-) q_minus = difference(q_minus, q_plus) */
static inline void fb_subtract(const double* q_plus, const double* q_minus, double* out) {
  double R[9], p[3];
  fb_relative(q_minus, q_plus, R, p);
  fb_log6(R, p, out);
  for (int j = 0; j < FB_NU; ++j) out[6 + j] = q_plus[7 + j] - q_minus[7 + j];
}
/* dSubtractdConfigurationPlus: d(q_plus (-) q_minus)/d q_plus = Jlog6(M) on the base block, +Id on the joints.
 * Only the 6x6 base block is returned (the joint block is +-Id and handled by the callers). */
static inline void fb_dsubtract_dplus(const double* q_plus, const double* q_minus, double* J6) {
  double R[9], p[3];
  fb_relative(q_minus, q_plus, R, p);
  fb_Jlog6(R, p, J6);
}
/* dSubtractdConfigurationMinus: base block -Jlog6(M) Ad(M^-1), -Id on the joints */
static inline void fb_dsubtract_dminus(const double* q_plus, const double* q_minus, double* J6) {
  double R[9], p[3], J1[36], X[36];
  fb_relative(q_minus, q_plus, R, p);
  fb_Jlog6(R, p, J1);
  /* X = -Ad(M^-1) = [[-R^T, R^T [p]x], [0, -R^T]] */
  double Sk[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0}, RtS[9];
  fb_mulT33(R, Sk, RtS);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      X[6 * r + k] = -R[3 * k + r];
      X[6 * r + 3 + k] = RtS[3 * r + k];
      X[6 * (3 + r) + k] = 0.0;
      X[6 * (3 + r) + 3 + k] = -R[3 * k + r];
    }
  for (int r = 0; r < 6; ++r)
    for (int k = 0; k < 6; ++k) {
      double acc = J1[6 * r] * X[k];
      for (int j = 1; j < 6; ++j) acc = fma(J1[6 * r + j], X[6 * j + k], acc);
      J6[6 * r + k] = acc;
    }
}
/* Robot::dSubtractdConfigurationInverse (robot.hxx:156-170): inverse of the block-upper-triangular 6x6
 * [[A, B], [0, D]] through the two 3x3 inverses (Eigen's closed-form cofactor inverse). */
static inline void fb_inv33(const double* A, double* Ai) {
  const double c00 = fma(A[4], A[8], -(A[5] * A[7]));
  const double c10 = fma(A[5], A[6], -(A[3] * A[8]));
  const double c20 = fma(A[3], A[7], -(A[4] * A[6]));
  const double det = fma(A[2], c20, fma(A[1], c10, A[0] * c00));
  const double id = 1.0 / det;
  Ai[0] = c00 * id; Ai[3] = c10 * id; Ai[6] = c20 * id;
  Ai[1] = fma(A[2], A[7], -(A[1] * A[8])) * id;
  Ai[4] = fma(A[0], A[8], -(A[2] * A[6])) * id;
  Ai[7] = fma(A[1], A[6], -(A[0] * A[7])) * id;
  Ai[2] = fma(A[1], A[5], -(A[2] * A[4])) * id;
  Ai[5] = fma(A[2], A[3], -(A[0] * A[5])) * id;
  Ai[8] = fma(A[0], A[4], -(A[1] * A[3])) * id;
}
static inline void fb_dsubtract_inverse(const double* J6, double* Jinv) {
  double A[9], B[9], D[9], Ai[9], Di[9], T[9], U[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) { A[3 * r + k] = J6[6 * r + k]; B[3 * r + k] = J6[6 * r + 3 + k]; D[3 * r + k] = J6[6 * (3 + r) + 3 + k]; }
  fb_inv33(A, Ai);
  fb_inv33(D, Di);
  fb_mul33(B, Di, T);
  fb_mul33(Ai, T, U);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      Jinv[6 * r + k] = Ai[3 * r + k];
      Jinv[6 * r + 3 + k] = -U[3 * r + k];
      Jinv[6 * (3 + r) + k] = 0.0;
      Jinv[6 * (3 + r) + 3 + k] = Di[3 * r + k];
    }
}
/* dIntegratedConfiguration (ARG0) = Ad(exp6(v)^-1) and dIntegratedVelocity (ARG1) = Jexp6(v), base blocks */
static inline void fb_dintegrate_dq(const double* v, double* J6) {
  double R[9], p[3];
  fb_exp6(v, R, p);
  double Sk[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0}, RtS[9];
  fb_mulT33(R, Sk, RtS);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      J6[6 * r + k] = R[3 * k + r];
      J6[6 * r + 3 + k] = -RtS[3 * r + k];
      J6[6 * (3 + r) + k] = 0.0;
      J6[6 * (3 + r) + 3 + k] = R[3 * k + r];
    }
}
static inline void fb_dintegrate_dv(const double* v, double* J6) { fb_Jexp6(v, J6); }

/* ---------------------------------------------------------------------------------------------- */
/* kinematics                                                                                      */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  double R[FB_NB][9], p[FB_NB][3]; /* world placement of every joint frame (body 0 = base) */
  double S[FB_NV][6];              /* world-frame motion subspace column of every dof */
  double ov[FB_NB][6];             /* spatial velocity of every body, world frame */
  double oa[FB_NB][6];             /* spatial acceleration WITHOUT gravity */
  double dV[FB_NV][6];             /* ov[parent] x S_c  (= pinocchio's dVdq column; zero for the base dofs) */
} fb_kin_t;

/* forwardKinematics(q, v, a) in the world frame.  v or a may be NULL (treated as zero). */
static inline void fb_forward_kinematics(const double* q, const double* v, const double* a, fb_kin_t* k) {
  fb_quat_to_R(q + 3, k->R[0]);
  for (int i = 0; i < 3; ++i) k->p[0][i] = q[i];
  for (int c = 0; c < 3; ++c) {
    const double e[3] = {k->R[0][c], k->R[0][3 + c], k->R[0][6 + c]};
    for (int i = 0; i < 3; ++i) { k->S[c][i] = e[i]; k->S[c][3 + i] = 0.0; }
    fb_cross(k->p[0], e, k->S[3 + c]);
    for (int i = 0; i < 3; ++i) k->S[3 + c][3 + i] = e[i];
  }
  for (int i = 0; i < 6; ++i) {
    double accv = 0.0, acca = 0.0;
    if (v) { accv = k->S[0][i] * v[0]; for (int c = 1; c < 6; ++c) accv = fma(k->S[c][i], v[c], accv); }
    if (a) { acca = k->S[0][i] * a[0]; for (int c = 1; c < 6; ++c) acca = fma(k->S[c][i], a[c], acca); }
    k->ov[0][i] = accv;
    k->oa[0][i] = acca;
  }
  for (int c = 0; c < 6; ++c)
    for (int i = 0; i < 6; ++i) k->dV[c][i] = 0.0;
  for (int j = 0; j < FB_NU; ++j) {
    const int b = 1 + j, pb = fb_parent_body(b), c = 6 + j;
    const double* Rp = k->R[pb];
    double* Rb = k->R[b];
    double sn, cs;
    oracle_canon_sincos(q[7 + j], &sn, &cs);
    if (ANYMAL_JOINT_AXIS[j] == 0) {       /* Rp * Rx */
      for (int i = 0; i < 3; ++i) {
        Rb[3 * i] = Rp[3 * i];
        Rb[3 * i + 1] = fma(cs, Rp[3 * i + 1], sn * Rp[3 * i + 2]);
        Rb[3 * i + 2] = fma(cs, Rp[3 * i + 2], -(sn * Rp[3 * i + 1]));
      }
    } else {                               /* Rp * Ry */
      for (int i = 0; i < 3; ++i) {
        Rb[3 * i] = fma(cs, Rp[3 * i], -(sn * Rp[3 * i + 2]));
        Rb[3 * i + 1] = Rp[3 * i + 1];
        Rb[3 * i + 2] = fma(cs, Rp[3 * i + 2], sn * Rp[3 * i]);
      }
    }
    const double* P = ANYMAL_JOINT_P[j];
    for (int i = 0; i < 3; ++i)
      k->p[b][i] = fma(Rp[3 * i + 2], P[2], fma(Rp[3 * i + 1], P[1], fma(Rp[3 * i], P[0], k->p[pb][i])));
    const int ax = ANYMAL_JOINT_AXIS[j];
    const double e[3] = {Rb[ax], Rb[3 + ax], Rb[6 + ax]};
    fb_cross(k->p[b], e, k->S[c]);
    for (int i = 0; i < 3; ++i) k->S[c][3 + i] = e[i];
    fb_mxm(k->ov[pb], k->S[c], k->dV[c]);
    const double qd = v ? v[c] : 0.0, qdd = a ? a[c] : 0.0;
    for (int i = 0; i < 6; ++i) {
      k->ov[b][i] = fma(k->S[c][i], qd, k->ov[pb][i]);
      k->oa[b][i] = fma(k->dV[c][i], qd, fma(k->S[c][i], qdd, k->oa[pb][i]));
    }
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* inverse dynamics and its derivatives                                                            */
/* ---------------------------------------------------------------------------------------------- */
/* world-frame spatial inertia of body b */
static inline void fb_body_inertia(const fb_kin_t* k, int b, fb_inertia_t* Y) {
  const double m = ANYMAL_MASS[b];
  double c[3], T[9];
  const double* R = k->R[b];
  for (int i = 0; i < 3; ++i)
    c[i] = fma(R[3 * i + 2], ANYMAL_COM[b][2], fma(R[3 * i + 1], ANYMAL_COM[b][1], fma(R[3 * i], ANYMAL_COM[b][0], k->p[b][i])));
  const double* Ic = ANYMAL_INERTIA[b];
  const double If[9] = {Ic[0], Ic[1], Ic[2], Ic[1], Ic[3], Ic[4], Ic[2], Ic[4], Ic[5]};
  fb_mul33(R, If, T);
  const double cc = fb_dot3(c, c);
  const int ii[6] = {0, 0, 0, 1, 1, 2}, jj[6] = {0, 1, 2, 1, 2, 2};
  for (int e = 0; e < 6; ++e) {
    const int i = ii[e], j = jj[e];
    const double iw = fma(T[3 * i + 2], R[3 * j + 2], fma(T[3 * i + 1], R[3 * j + 1], T[3 * i] * R[3 * j]));
    const double par = (i == j) ? (cc - c[i] * c[j]) : -(c[i] * c[j]);
    Y->I[e] = fma(m, par, iw);
  }
  Y->m = m;
  for (int i = 0; i < 3; ++i) Y->h[i] = m * c[i];
}

/* world-frame spatial force (about the world origin) of contact i carrying the LOCAL force f (frame axes =
 * axes of the parent joint frame): Robot::setContactForces, point_contact.hxx:15-20 (jXf.act(Force(f, 0))) */
static inline void fb_contact_point(const fb_kin_t* k, int i, double* P) {
  const int b = 1 + ANYMAL_CONTACT_PARENT_JOINT[i];
  const double* R = k->R[b];
  const double* pc = ANYMAL_CONTACT_P[i];
  for (int r = 0; r < 3; ++r) P[r] = fma(R[3 * r + 2], pc[2], fma(R[3 * r + 1], pc[1], fma(R[3 * r], pc[0], k->p[b][r])));
}
static inline void fb_contact_wrench(const fb_kin_t* k, int i, const double* f, double* W) {
  const int b = 1 + ANYMAL_CONTACT_PARENT_JOINT[i];
  double P[3];
  fb_contact_point(k, i, P);
  fb_rot(k->R[b], f, W);
  fb_cross(P, W, W + 3);
}

typedef struct {
  double U[FB_NV][6];    /* Ycrb S_c   (= dFda column) */
  double F[FB_NB][6];    /* composite force of the subtree of every body */
} fb_dyn_t;

/* RNEA + computeRNEADerivatives with external forces (robot.hxx:444-500).  gravity = 9.81 for the
 * continuous dynamics, 0 for RNEAImpulse (robot.cpp: impulse_model_.gravity = 0).  f[4][3] are the LOCAL
 * contact forces (zero rows for inactive contacts).  tau always; dq, dv, M (row-major 18x18) when non-NULL
 * (dv may be NULL alone: RNEAImpulseDerivatives discards it).  The caller has run
 * fb_forward_kinematics(q, v, a). */
static inline void fb_rnea_derivatives(const fb_kin_t* k, const double f[FB_NC][3], double gravity, double* tau,
                                       double* dq, double* dv, double* M) {
  fb_inertia_t Y[FB_NB];
  fb_dinertia_t D[FB_NB];
  double F[FB_NB][6], agf[FB_NB][6];
  for (int b = 0; b < FB_NB; ++b) {
    fb_body_inertia(k, b, &Y[b]);
    for (int i = 0; i < 6; ++i) agf[b][i] = k->oa[b][i];
    agf[b][2] = k->oa[b][2] + gravity;
    double Ya[6], h[6], vh[6];
    fb_Ymul(&Y[b], agf[b], Ya);
    fb_Ymul(&Y[b], k->ov[b], h);
    fb_mxf(k->ov[b], h, vh);
    for (int i = 0; i < 6; ++i) F[b][i] = Ya[i] + vh[i];
    fb_dinertia(&Y[b], k->ov[b], &D[b]);
  }
  for (int i = 0; i < FB_NC; ++i) {
    const int b = 1 + ANYMAL_CONTACT_PARENT_JOINT[i];
    double W[6];
    fb_contact_wrench(k, i, f[i], W);
    for (int e = 0; e < 6; ++e) F[b][e] -= W[e];
  }
  /* composites: leaves to root, each body into its parent (pinocchio's backward pass order) */
  for (int b = FB_NB - 1; b >= 1; --b) {
    const int pb = fb_parent_body(b);
    Y[pb].m += Y[b].m;
    for (int i = 0; i < 3; ++i) { Y[pb].h[i] += Y[b].h[i]; D[pb].pl[i] += D[b].pl[i]; D[pb].pa[i] += D[b].pa[i]; }
    for (int i = 0; i < 6; ++i) { Y[pb].I[i] += Y[b].I[i]; D[pb].S[i] += D[b].S[i]; F[pb][i] += F[b][i]; }
  }
  for (int c = 0; c < FB_NV; ++c) tau[c] = fb_dot6(k->S[c], F[fb_body_of_dof(c)]);
  if (!M) return;
  double U[FB_NV][6], W[FB_NV][6], dFv[FB_NV][6], dFq[FB_NV][6], dFqa[FB_NV][6], dAq[FB_NV][6], dAv[FB_NV][6];
  const double a0[6] = {0.0, 0.0, gravity, 0.0, 0.0, 0.0};
  for (int c = 0; c < FB_NV; ++c) {
    const int b = fb_body_of_dof(c);
    double dJ[6], t1[6], t2[6];
    fb_mxm(k->ov[b], k->S[c], dJ);
    if (b == 0) {
      fb_mxm(a0, k->S[c], dAq[c]);
    } else {
      const int pb = fb_parent_body(b);
      fb_mxm(agf[pb], k->S[c], t1);
      fb_mxm(k->ov[pb], k->dV[c], t2);
      for (int i = 0; i < 6; ++i) dAq[c][i] = t1[i] + t2[i];
    }
    for (int i = 0; i < 6; ++i) dAv[c][i] = dJ[i] + k->dV[c][i];
    fb_Ymul(&Y[b], k->S[c], U[c]);
    fb_DTmul(&D[b], k->S[c], W[c]);
    fb_Dmul(&D[b], k->S[c], t1);
    fb_Ymul(&Y[b], dAv[c], t2);
    for (int i = 0; i < 6; ++i) dFv[c][i] = t1[i] + t2[i];
    fb_Dmul(&D[b], k->dV[c], t1);
    fb_Ymul(&Y[b], dAq[c], t2);
    for (int i = 0; i < 6; ++i) dFq[c][i] = t1[i] + t2[i];
    fb_mxf(k->S[c], F[b], t1);
    for (int i = 0; i < 6; ++i) dFqa[c][i] = dFq[c][i] + t1[i];
  }
  for (int r = 0; r < FB_NV; ++r)
    for (int c = 0; c < FB_NV; ++c) {
      double eq = 0.0, ev = 0.0, em = 0.0;
      if (fb_same_joint(r, c)) {
        eq = fb_dot6(k->S[r], dFq[c]);
        ev = fb_dot6(k->S[r], dFv[c]);
        em = fb_dot6(k->S[r], U[c]);
      } else if (fb_is_ancestor(r, c)) {
        eq = fb_dot6(k->S[r], dFqa[c]);
        ev = fb_dot6(k->S[r], dFv[c]);
        em = fb_dot6(k->S[r], U[c]);
      } else if (fb_is_ancestor(c, r)) {
        eq = fb_dot6(dAq[c], U[r]) + fb_dot6(k->dV[c], W[r]);
        ev = fb_dot6(dAv[c], U[r]) + fb_dot6(k->S[c], W[r]);
      }
      if (dq) dq[r * FB_NV + c] = eq;
      if (dv) dv[r * FB_NV + c] = ev;
      M[r * FB_NV + c] = em;
    }
  /* robot.hxx:496-499: strictly lower triangle of dRNEA/da := transpose of the upper one */
  for (int r = 1; r < FB_NV; ++r)
    for (int c = 0; c < r; ++c) M[r * FB_NV + c] = M[c * FB_NV + r];
}

/* ---------------------------------------------------------------------------------------------- */
/* point contacts (robot/point_contact.hxx)                                                        */
/* ---------------------------------------------------------------------------------------------- */
/* y = oMf^-1 x for a world-frame motion vector x and the contact frame (R_f, P_f) */
static inline void fb_pullback(const double* Rf, const double* Pf, const double* x, double* y) {
  double t[3], u[3];
  fb_cross(x + 3, Pf, t);
  for (int i = 0; i < 3; ++i) u[i] = x[i] + t[i];
  fb_rotT(Rf, u, y);
  fb_rotT(Rf, x + 3, y + 3);
}
static inline int fb_in_support(int contact, int c) { return c < 6 || (c - 6) / 3 == contact; }

typedef struct {
  double P[3];              /* world position of the contact frame */
  double vF[6], aF[6];      /* LOCAL frame velocity / spatial acceleration */
  double J[6][FB_NV];       /* getFrameJacobian(LOCAL) = d vF / d v = d aF / d a */
  double v_dq[6][FB_NV];    /* getFrameVelocityDerivatives: d vF / d q */
  double a_dq[6][FB_NV];    /* getFrameAccelerationDerivatives: d aF / d q */
  double a_dv[6][FB_NV];    /*                                   d aF / d v */
} fb_frame_t;

/* LOCAL frame kinematics and derivatives of contact frame i (pinocchio frames-derivatives.hpp), from the
 * world-frame recursion quantities:
 *   d vF/d q_c = oMf^-1 (ov_p x S_c)                              (zero for the base dofs)
 *   d aF/d v_c = oMf^-1 (ov_b x S_c + S_c x (ov_i - ov_p))
 *   d aF/d q_c = oMf^-1 (oa_p x S_c + (ov_p x S_c) x (ov_i - ov_p))
 * with b the body of dof c, p its parent (universe: zero), i the body carrying the frame, oa without gravity. */
static inline void fb_frame_kinematics(const fb_kin_t* k, int contact, int level, fb_frame_t* fr) {
  const int bi = 1 + ANYMAL_CONTACT_PARENT_JOINT[contact];
  const double* Rf = k->R[bi];
  fb_contact_point(k, contact, fr->P);
  memset(fr->J, 0, sizeof(fr->J));
  memset(fr->v_dq, 0, sizeof(fr->v_dq));
  memset(fr->a_dq, 0, sizeof(fr->a_dq));
  memset(fr->a_dv, 0, sizeof(fr->a_dv));
  fb_pullback(Rf, fr->P, k->ov[bi], fr->vF);
  fb_pullback(Rf, fr->P, k->oa[bi], fr->aF);
  for (int c = 0; c < FB_NV; ++c) {
    if (!fb_in_support(contact, c)) continue;
    double y[6];
    fb_pullback(Rf, fr->P, k->S[c], y);
    for (int e = 0; e < 6; ++e) fr->J[e][c] = y[e];
    if (level < 1) continue;
    const int b = fb_body_of_dof(c);
    double u[6], x[6], t1[6], t2[6];
    if (b == 0) {
      for (int e = 0; e < 6; ++e) u[e] = k->ov[bi][e];
    } else {
      const int pb = fb_parent_body(b);
      for (int e = 0; e < 6; ++e) u[e] = k->ov[bi][e] - k->ov[pb][e];
      fb_pullback(Rf, fr->P, k->dV[c], y);
      for (int e = 0; e < 6; ++e) fr->v_dq[e][c] = y[e];
    }
    if (level < 2) continue;
    fb_mxm(k->ov[b], k->S[c], t1);
    fb_mxm(k->S[c], u, t2);
    for (int e = 0; e < 6; ++e) x[e] = t1[e] + t2[e];
    fb_pullback(Rf, fr->P, x, y);
    for (int e = 0; e < 6; ++e) fr->a_dv[e][c] = y[e];
    if (b != 0) {
      const int pb = fb_parent_body(b);
      fb_mxm(k->oa[pb], k->S[c], t1);
      fb_mxm(k->dV[c], u, t2);
      for (int e = 0; e < 6; ++e) x[e] = t1[e] + t2[e];
      fb_pullback(Rf, fr->P, x, y);
      for (int e = 0; e < 6; ++e) fr->a_dq[e][c] = y[e];
    }
  }
}

/* PointContact::computeBaumgarteResidual (point_contact.hxx:67-86): classical acceleration + 2/D velocity +
 * 1/D^2 (world position - contact point); D = baumgarte_time_step. */
static inline void fb_baumgarte_residual(const fb_frame_t* fr, double time_step, const double* contact_point, double* C) {
  const double wv = 2.0 / time_step, wp = 1.0 / (time_step * time_step);
  double wxv[3];
  fb_cross(fr->vF + 3, fr->vF, wxv);
  for (int i = 0; i < 3; ++i) {
    const double acl = fr->aF[i] + wxv[i];
    C[i] = fma(wp, fr->P[i] - contact_point[i], fma(wv, fr->vF[i], acl));
  }
}
/* PointContact::computeBaumgarteDerivatives (point_contact.hxx:89-144), accumulation order as written there.
 * NOTE the reference adds skew(v_lin) d(omega) where the exact derivative of omega x v has the opposite sign;
 * kept as is (the Newton iteration only needs a consistent residual), see DESIGN.md. dC*: 3 x 18 row-major. */
static inline void fb_baumgarte_derivatives(const fb_kin_t* k, int contact, const fb_frame_t* fr, double time_step,
                                            double* dCdq, double* dCdv, double* dCda) {
  const int bi = 1 + ANYMAL_CONTACT_PARENT_JOINT[contact];
  const double* Rf = k->R[bi];
  const double wv = 2.0 / time_step, wp = 1.0 / (time_step * time_step);
  const double* vl = fr->vF;
  const double* va = fr->vF + 3;
  for (int c = 0; c < FB_NV; ++c) {
    double q3[3], v3[3];
    const double vq_l[3] = {fr->v_dq[0][c], fr->v_dq[1][c], fr->v_dq[2][c]}, vq_a[3] = {fr->v_dq[3][c], fr->v_dq[4][c], fr->v_dq[5][c]};
    const double J_l[3] = {fr->J[0][c], fr->J[1][c], fr->J[2][c]}, J_a[3] = {fr->J[3][c], fr->J[4][c], fr->J[5][c]};
    double t1[3], t2[3], RJ[3];
    fb_cross(va, vq_l, t1);
    fb_cross(vl, vq_a, t2);
    fb_rot(Rf, J_l, RJ);
    for (int i = 0; i < 3; ++i) q3[i] = fma(wp, RJ[i], fma(wv, vq_l[i], (fr->a_dq[i][c] + t1[i]) + t2[i]));
    fb_cross(va, J_l, t1);
    fb_cross(vl, J_a, t2);
    for (int i = 0; i < 3; ++i) v3[i] = fma(wv, J_l[i], (fr->a_dv[i][c] + t1[i]) + t2[i]);
    for (int i = 0; i < 3; ++i) {
      dCdq[i * FB_NV + c] = q3[i];
      dCdv[i * FB_NV + c] = v3[i];
      dCda[i * FB_NV + c] = J_l[i];
    }
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* dense symmetric helpers: Cholesky with reciprocal pivots (canonical form of oracle/idocp_oracle.c) */
/* ---------------------------------------------------------------------------------------------- */
/* A (n x n, leading dimension lda, lower triangle read) = L L^T; L stored "L[j*ldl + i]" = L_ij, i >= j
 * (i.e. column j contiguous); rd[k] = 1 / L_kk.  Returns 0 or k+1 for the first non-positive pivot. */
static inline int fb_llt(const double* A, int lda, int n, double* L, int ldl, double* rd) {
  int info = 0;
  for (int k = 0; k < n; ++k) {
    double x = A[k * lda + k];
    for (int j = 0; j < k; ++j) x = fma(-L[j * ldl + k], L[j * ldl + k], x);
    if (!canon_pivot_ok(x) && !info) info = k + 1;
    rd[k] = canon_rsqrt(x);
    L[k * ldl + k] = x * rd[k];
    for (int i = k + 1; i < n; ++i) {
      double y = A[i * lda + k];
      for (int j = 0; j < k; ++j) y = fma(-L[j * ldl + i], L[j * ldl + k], y);
      L[k * ldl + i] = y * rd[k];
    }
  }
  return info;
}
/* x := (L L^T)^-1 x for one right-hand side with stride incx */
static inline void fb_llt_solve(const double* L, int ldl, const double* rd, int n, double* x, int incx) {
  for (int i = 0; i < n; ++i) {
    double y = x[i * incx];
    for (int j = 0; j < i; ++j) y = fma(-L[j * ldl + i], x[j * incx], y);
    x[i * incx] = y * rd[i];
  }
  for (int i = n - 1; i >= 0; --i) {   /* L^T x = z: the terms of row i are subtracted from the last column backwards */
    double y = x[i * incx];
    for (int j = n - 1; j > i; --j) y = fma(-L[i * ldl + j], x[j * incx], y);
    x[i * incx] = y * rd[i];
  }
}

/* Robot::computeMJtJinv (robot.hxx:576-615): MJtJinv = [[M, J^T], [J, 0]]^-1, (18+dimf)^2 row-major with
 * leading dimension ld.  The reference goes through pinocchio's sparse U D U^T of M, which eliminates the joints from the
 * leaves to the root so that the tree sparsity of M survives (a leg joint couples only with its own leg and the base).  Here:
 * a Cholesky of M in the same kind of order -- FB_MPERM: joint s of leg i at position 4 s + i, the base last -- written as the
 * plain dense factorisation of the permuted matrix (the structurally zero terms contribute exact zeros, which is what lets the
 * kernel skip them and share one step between the four legs); info counts pivots in that order.  Blocks in the reference's order:
 *   Minv = M^-1;  S = J Minv J^T;  BR = -S^-1;  BL = J Minv;  TR = BL^T (-BR);  TL = Minv - TR BL;  BL = TR^T */
static const int FB_MPERM[FB_NV] = {6, 9, 12, 15, 7, 10, 13, 16, 8, 11, 14, 17, 0, 1, 2, 3, 4, 5};
static inline int fb_MJtJinv(const double* M, const double* J, int dimf, double* out, int ld) {
  const int n = FB_NV;
  double L[FB_NV * FB_NV], rd[FB_NV], Minv[FB_NV * FB_NV], JMi[FB_MAXF * FB_NV], S[FB_MAXF * FB_MAXF], Ls[FB_MAXF * FB_MAXF],
      rds[FB_MAXF], Si[FB_MAXF * FB_MAXF], Mp[FB_NV * FB_NV], Xp[FB_NV * FB_NV];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) Mp[r * n + c] = M[FB_MPERM[r] * n + FB_MPERM[c]];
  int info = fb_llt(Mp, n, n, L, n, rd);
  for (int c = 0; c < n; ++c) {
    for (int r = 0; r < n; ++r) Xp[r * n + c] = (r == c) ? 1.0 : 0.0;
    fb_llt_solve(L, n, rd, n, Xp + c, n);
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) Minv[FB_MPERM[r] * n + FB_MPERM[c]] = Xp[r * n + c];
  for (int r = 0; r < dimf; ++r)
    for (int c = 0; c < n; ++c) {
      double acc = J[r * n] * Minv[c];
      for (int j = 1; j < n; ++j) acc = fma(J[r * n + j], Minv[j * n + c], acc);
      JMi[r * n + c] = acc;
    }
  for (int r = 0; r < dimf; ++r)
    for (int c = 0; c < dimf; ++c) {
      double acc = JMi[r * n] * J[c * n];
      for (int j = 1; j < n; ++j) acc = fma(JMi[r * n + j], J[c * n + j], acc);
      S[r * dimf + c] = acc;
    }
  if (dimf > 0) {
    const int i2 = fb_llt(S, dimf, dimf, Ls, dimf, rds);
    if (i2 && !info) info = 100 + i2;
  }
  for (int c = 0; c < dimf; ++c) {
    for (int r = 0; r < dimf; ++r) Si[r * dimf + c] = (r == c) ? 1.0 : 0.0;
    fb_llt_solve(Ls, dimf, rds, dimf, Si + c, dimf);
  }
  /* BR = -S^-1 */
  for (int r = 0; r < dimf; ++r)
    for (int c = 0; c < dimf; ++c) out[(n + r) * ld + n + c] = -Si[r * dimf + c];
  /* TR = (J Minv)^T S^-1 */
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < dimf; ++c) {
      double acc = 0.0;
      for (int j = 0; j < dimf; ++j) acc = (j == 0) ? JMi[j * n + r] * Si[j * dimf + c] : fma(JMi[j * n + r], Si[j * dimf + c], acc);
      out[r * ld + n + c] = acc;
    }
  /* TL = Minv - TR (J Minv) */
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      double acc = Minv[r * n + c];
      for (int j = 0; j < dimf; ++j) acc = fma(-out[r * ld + n + j], JMi[j * n + c], acc);
      out[r * ld + c] = acc;
    }
  /* BL = TR^T */
  for (int r = 0; r < dimf; ++r)
    for (int c = 0; c < n; ++c) out[(n + r) * ld + c] = out[c * ld + n + r];
  return info;
}

#endif /* ORACLE_FB_ROBOT_H_ */

// fb_math.cuh -- scalar building blocks of the floating-base (ANYmal) kernels: 3-vectors, spatial motion / force
// algebra in the world frame, spatial inertias and their velocity derivative ("doYcrb"), SO(3)/SE(3) exp / log and
// their Jacobians, and the configuration-space operators of the free-flyer + revolute tree.
//
// These replace what the reference obtains from pinocchio (absent from /root/reference) at the call sites
//   Robot::integrateConfiguration / subtractConfiguration / dSubtractdConfiguration* / dIntegrated*
//                                                      include/idocp/robot/robot.hxx:22-170
//   Robot::RNEA / RNEADerivatives / RNEAImpulse*       robot.hxx:444-535 (pinocchio::rnea, computeRNEADerivatives)
// restated from the published algorithms (explog.hpp, liegroup/special-euclidean.hpp, rnea-derivatives.hxx).
// Conventions: q = [p, quaternion xyzw, 12 joint angles], v = [base linear, base angular (base frame), joint rates];
// spatial vectors [linear; angular] in the WORLD frame; 3x3 / 6x6 matrices row-major.
// CANONICAL ARITHMETIC: -fmad=false, every fused multiply-add is an explicit fma(), sin/cos/acos are
// the shared polynomial implementations of octet.cuh -- the kernels' results are reproducible bit for bit.
#pragma once
#include "octet.cuh"

#ifdef IDOCP_B200_EMU
#define ANYMAL_TABLE static const
#else
#define ANYMAL_TABLE __device__ const
#endif
#include "model_anymal.h"

namespace idocp_b200 {

#define FB_NV 18
#define FB_NQ 19
#define FB_NU 12
#define FB_NB 13
#define FB_NC 4
#define FB_MAXF 12
#define FB_NX 36
#define FB_NVF 30
#define FB_NPASS 6
#define FB_TAYLOR 1.220703125e-04 /* TaylorSeriesExpansion<double>::precision<3>() = 2^-13 */
#define FB_PI 3.14159265358979311600e+00

/* ---------------------------------------------------------------------------------------------- */
/* small vectors                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
__device__ inline void fb_cross(const double* a, const double* b, double* c) { /* c must not alias a, b */
  c[0] = fma(a[1], b[2], -(a[2] * b[1]));
  c[1] = fma(a[2], b[0], -(a[0] * b[2]));
  c[2] = fma(a[0], b[1], -(a[1] * b[0]));
}
__device__ inline double fb_dot3(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
/* y = R x  /  y = R^T x, R row-major 3x3, y must not alias x */
__device__ inline void fb_rot(const double* R, const double* x, double* y) {
  for (int i = 0; i < 3; ++i) y[i] = fma(R[3 * i + 2], x[2], fma(R[3 * i + 1], x[1], R[3 * i] * x[0]));
}
__device__ inline void fb_rotT(const double* R, const double* x, double* y) {
  for (int i = 0; i < 3; ++i) y[i] = fma(R[6 + i], x[2], fma(R[3 + i], x[1], R[i] * x[0]));
}
/* C = A B, 3x3 row-major */
__device__ inline void fb_mul33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(A[3 * r + 2], B[6 + k], fma(A[3 * r + 1], B[3 + k], A[3 * r] * B[k]));
}
/* C = A^T B */
__device__ inline void fb_mulT33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(A[6 + r], B[6 + k], fma(A[3 + r], B[3 + k], A[r] * B[k]));
}
/* M += skew(v) */
__device__ inline void fb_add_skew(const double* v, double* M) {
  M[1] -= v[2]; M[2] += v[1]; M[3] += v[2]; M[5] -= v[0]; M[6] -= v[1]; M[7] += v[0];
}
/* y = A x, A symmetric stored (xx,xy,xz,yy,yz,zz) */
__device__ inline void fb_sym3(const double* A, const double* x, double* y) {
  y[0] = fma(A[2], x[2], fma(A[1], x[1], A[0] * x[0]));
  y[1] = fma(A[4], x[2], fma(A[3], x[1], A[1] * x[0]));
  y[2] = fma(A[5], x[2], fma(A[4], x[1], A[2] * x[0]));
}

/* spatial motion cross product c = a x b; force cross c = a x* f; pairing <m, f> */
__device__ inline void fb_mxm(const double* a, const double* b, double* c) {
  double t1[3], t2[3];
  fb_cross(a + 3, b, t1);
  fb_cross(a, b + 3, t2);
  for (int i = 0; i < 3; ++i) c[i] = t1[i] + t2[i];
  fb_cross(a + 3, b + 3, c + 3);
}
__device__ inline void fb_mxf(const double* a, const double* f, double* c) {
  double t1[3], t2[3];
  fb_cross(a + 3, f, c);
  fb_cross(a + 3, f + 3, t1);
  fb_cross(a, f, t2);
  for (int i = 0; i < 3; ++i) c[3 + i] = t1[i] + t2[i];
}
__device__ inline double fb_dot6(const double* m, const double* f) {
  double acc = m[0] * f[0];
  for (int i = 1; i < 6; ++i) acc = fma(m[i], f[i], acc);
  return acc;
}

/* spatial inertia about the world origin: mass, first moment h = m c, rotational inertia about the origin */
typedef struct { double m, h[3], I[6]; } fb_inertia_t;
/* f = Y mv : lin = m u - h x w, ang = I w + h x u */
__device__ inline void fb_Ymul(const fb_inertia_t* Y, const double* mv, double* f) {
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
__device__ inline void fb_dinertia(const fb_inertia_t* Y, const double* v, fb_dinertia_t* D) {
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
__device__ inline void fb_Dmul(const fb_dinertia_t* D, const double* m, double* y) {
  double t[3], s[3], u[3];
  fb_cross(m + 3, D->pl, t);          /* w x pl */
  fb_sym3(D->S, m + 3, s);
  fb_cross(D->pa, m + 3, u);          /* pa x w */
  for (int i = 0; i < 3; ++i) { y[i] = 2.0 * t[i]; y[3 + i] = s[i] - u[i]; }
}
__device__ inline void fb_DTmul(const fb_dinertia_t* D, const double* m, double* y) {
  double t[3], s[3], u[3];
  fb_cross(D->pl, m, t);              /* pl x u */
  fb_sym3(D->S, m + 3, s);
  fb_cross(D->pa, m + 3, u);          /* pa x w */
  for (int i = 0; i < 3; ++i) { y[i] = 0.0; y[3 + i] = fma(2.0, t[i], s[i]) + u[i]; }
}

/* ---------------------------------------------------------------------------------------------- */
/* tree structure                                                                                  */
/* ---------------------------------------------------------------------------------------------- */
__device__ inline int fb_body_of_dof(int c) { return c < 6 ? 0 : c - 5; }
__device__ inline int fb_parent_body(int b) { return ANYMAL_JOINT_PARENT[b - 1] + 1; } /* b >= 1 */
/* 1 when dof r belongs to a strict ancestor joint of dof c */
__device__ inline int fb_is_ancestor(int r, int c) {
  if (r < 6) return c >= 6;
  if (c < 6) return 0;
  return ((r - 6) / 3 == (c - 6) / 3) && r < c;
}
__device__ inline int fb_same_joint(int r, int c) { return (r < 6 && c < 6) || r == c; }

/* ---------------------------------------------------------------------------------------------- */
/* SO(3) / SE(3): pinocchio::exp3/log3/Jlog3/exp6/log6/Jlog6/Jexp6 (explog.hpp), Eigen quaternions      */
/* ---------------------------------------------------------------------------------------------- */
/* Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w) */
__device__ inline void fb_quat_to_R(const double* q, double* R) {
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
__device__ inline void fb_R_to_quat(const double* R, double* q) {
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

__device__ inline void fb_log3(const double* R, double* theta_out, double* w) {
  const double tr = (R[0] + R[4]) + R[8];
  double theta;
  if (tr > 3.0) theta = 0.0;
  else if (tr < -1.0) theta = FB_PI;
  else theta = canon_acos((tr - 1.0) * 0.5);
  *theta_out = theta;
  if (theta >= FB_PI - 1e-2) {
    double sn, cphi;
    canon_sincos(theta - FB_PI, &sn, &cphi);
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
    canon_sincos(theta, &sn, &cs);
    t = theta / sn;
  }
  t *= 0.5;
  w[0] = t * (R[7] - R[5]); w[1] = t * (R[2] - R[6]); w[2] = t * (R[3] - R[1]);
}
__device__ inline void fb_Jlog3(double theta, const double* w, double* A) {
  double alpha, diag;
  if (theta < FB_TAYLOR) {
    alpha = 1.0 / 12.0 + (theta * theta) / 720.0;
    diag = 0.5 * (2.0 - (theta * theta) / 6.0);
  } else {
    double st, ct;
    canon_sincos(theta, &st, &ct);
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
__device__ inline void fb_log6(const double* R, const double* p, double* out) {
  double theta, w[3];
  fb_log3(R, &theta, w);
  const double t2 = theta * theta;
  double alpha, beta;
  if (theta < FB_TAYLOR) {
    alpha = (1.0 - t2 / 12.0) - (t2 * t2) / 720.0;
    beta = 1.0 / 12.0 + t2 / 720.0;
  } else {
    double st, ct;
    canon_sincos(theta, &st, &ct);
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
__device__ inline void fb_se3_C(double theta, const double* w, const double* p, double* C) {
  const double t2 = theta * theta;
  double beta, bdot;
  if (theta < FB_TAYLOR) {
    beta = 1.0 / 12.0 + t2 / 720.0;
    bdot = 1.0 / 360.0;
  } else {
    double st, ct;
    canon_sincos(theta, &st, &ct);
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
__device__ inline void fb_Jlog6(const double* R, const double* p, double* J) {
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
__device__ inline void fb_exp6(const double* nu, double* R, double* p) {
  const double* v = nu;
  const double* w = nu + 3;
  const double t2 = fb_dot3(w, w);
  const double t = sqrt(t2);
  double alpha_wxv, alpha_v, alpha_w, diag;
  if (t > FB_TAYLOR) {
    double st, ct;
    canon_sincos(t, &st, &ct);
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
__device__ inline void fb_Jexp3(const double* r, double* J) {
  const double n2 = fb_dot3(r, r);
  const double n = sqrt(n2);
  double a, b, c;
  if (n < FB_TAYLOR) {
    a = 1.0 - n2 / 6.0;
    b = -0.5 - n2 / 24.0;
    c = 1.0 / 6.0 - n2 / 120.0;
  } else {
    double sn, cn;
    canon_sincos(n, &sn, &cn);
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
__device__ inline void fb_Jexp6(const double* nu, double* J) {
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
__device__ inline void fb_integrate(const double* q, const double* v, double alpha, double* q_out) {
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
__device__ inline void fb_relative(const double* q0, const double* q1, double* R, double* p) {
  double R0[9], R1[9], dp[3];
  fb_quat_to_R(q0 + 3, R0);
  fb_quat_to_R(q1 + 3, R1);
  fb_mulT33(R0, R1, R);
  for (int i = 0; i < 3; ++i) dp[i] = q1[i] - q0[i];
  fb_rotT(R0, dp, p);
}
/* Robot::subtractConfiguration(q_plus, q_minus, out): out = q_plus (-) q_minus = difference(q_minus, q_plus) */
__device__ inline void fb_subtract(const double* q_plus, const double* q_minus, double* out) {
  double R[9], p[3];
  fb_relative(q_minus, q_plus, R, p);
  fb_log6(R, p, out);
  for (int j = 0; j < FB_NU; ++j) out[6 + j] = q_plus[7 + j] - q_minus[7 + j];
}
/* dSubtractdConfigurationPlus: d(q_plus (-) q_minus)/d q_plus = Jlog6(M) on the base block, +Id on the joints.
 * Only the 6x6 base block is returned (the joint block is +-Id and handled by the callers). */
__device__ inline void fb_dsubtract_dplus(const double* q_plus, const double* q_minus, double* J6) {
  double R[9], p[3];
  fb_relative(q_minus, q_plus, R, p);
  fb_Jlog6(R, p, J6);
}
/* dSubtractdConfigurationMinus: base block -Jlog6(M) Ad(M^-1), -Id on the joints */
__device__ inline void fb_dsubtract_dminus(const double* q_plus, const double* q_minus, double* J6) {
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
__device__ inline void fb_inv33(const double* A, double* Ai) {
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
__device__ inline void fb_dsubtract_inverse(const double* J6, double* Jinv) {
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
__device__ inline void fb_dintegrate_dq(const double* v, double* J6) {
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
__device__ inline void fb_dintegrate_dv(const double* v, double* J6) { fb_Jexp6(v, J6); }


}  // namespace idocp_b200

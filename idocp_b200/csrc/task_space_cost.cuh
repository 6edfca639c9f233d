// task_space_cost.cuh -- TimeVaryingTaskSpace6DCost on the end-effector frame, lane-parallel.
//
// Reference: src/cost/time_varying_task_space_6d_cost.cpp:68-195 (cost, gradient JJ^T W diff, Gauss-Newton
// Hessian JJ^T W JJ with JJ = Jlog6(diff_SE3) * J_frame(LOCAL)), Robot::framePlacement / getFrameJacobian
// (robot.hxx:186,193-203).  pinocchio's log3 / log6 / Jlog3 / Jlog6 (spatial/log.hxx, absent here) are
// restated from their published formulas; the oracle's task_evaluate repeats every operation below.
//
// Every lane of the octet evaluates the (instance-wide) pose error and Jlog6 redundantly -- identical
// operations give identical bits in all lanes -- and owns column `lane` of the frame Jacobian and of JJ.
// The 6D reference of the stage's time is sampled on the host (the user's compute_q_6d_ref virtual) into a
// per-stage table [R row-major (9), p (3)].
#pragma once
#include "chain_dynamics.cuh"

namespace idocp_b200 {

constexpr double TASK_TAYLOR = 1.220703125e-04;  // TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4)
constexpr double TASK_PI = 3.14159265358979311600e+00;

__device__ __forceinline__ double dot3r(const double* row, V3 x) { return fma(row[2], x.z, fma(row[1], x.y, row[0] * x.x)); }
// y = M^T x for a row-major 3x3
__device__ __forceinline__ V3 mulT3(const double* M, V3 x) {
  return V3{fma(M[6], x.z, fma(M[3], x.y, M[0] * x.x)), fma(M[7], x.z, fma(M[4], x.y, M[1] * x.x)),
            fma(M[8], x.z, fma(M[5], x.y, M[2] * x.x))};
}
__device__ __forceinline__ void add_skew(V3 v, double* M) {
  M[1] -= v.z; M[2] += v.y; M[3] += v.z; M[5] -= v.x; M[6] -= v.y; M[7] += v.x;
}

// pinocchio::log3(R, theta)
__device__ __forceinline__ V3 log3_canon(const double* R, double& theta) {
  const double tr = (R[0] + R[4]) + R[8];
  if (tr > 3.0) theta = 0.0;
  else if (tr < -1.0) theta = TASK_PI;
  else theta = canon_acos((tr - 1.0) * 0.5);
  if (theta >= TASK_PI - 1e-2) {
    double sn, cphi;
    canon_sincos(theta - TASK_PI, &sn, &cphi);
    const double beta = (theta * theta) / (1.0 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    return V3{(R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0.0 ? sqrt(t0) : 0.0),
              (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0.0 ? sqrt(t1) : 0.0),
              (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0.0 ? sqrt(t2) : 0.0)};
  }
  double t = 1.0;
  if (theta > TASK_TAYLOR) {
    double sn, cs;
    canon_sincos(theta, &sn, &cs);
    t = theta / sn;
  }
  t *= 0.5;
  return V3{t * (R[7] - R[5]), t * (R[2] - R[6]), t * (R[3] - R[1])};
}

// pinocchio::Jlog3(theta, log, Jlog)
__device__ __forceinline__ void jlog3_canon(double theta, V3 w, double* A) {
  double alpha, diag;
  if (theta < TASK_TAYLOR) {
    alpha = 1.0 / 12.0 + (theta * theta) / 720.0;
    diag = 0.5 * (2.0 - (theta * theta) / 6.0);
  } else {
    double st, ct;
    canon_sincos(theta, &st, &ct);
    const double st_1mct = st / (1.0 - ct);
    alpha = 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
    diag = 0.5 * (theta * st_1mct);
  }
  const V3 aw = alpha * w;
  const double wv[3] = {w.x, w.y, w.z}, av[3] = {aw.x, aw.y, aw.z};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) A[3 * r + k] = av[r] * wv[k];
  A[0] += diag; A[4] += diag; A[8] += diag;
  add_skew(0.5 * w, A);
}

struct TaskEval {
  double diff[6];  // log6(SE3_ref^-1 * oMf) = [linear; angular]   (same in every lane)
  double JJ[6];    // column `lane` of Jlog6 * frame Jacobian (LOCAL)
};

// R, p: world placement of this lane's joint (chain_fk); ee: [R row-major (9), p (3)] placement of the frame in
// the last joint's frame; ref: the stage's SE3 reference.
// kind = DevProblem::task_enabled: 1 = the 6D cost above; 2 = TaskSpace3DCost / TimeVaryingTaskSpace3DCost
// (src/cost/task_space_3d_cost.cpp:56-140): diff_3d = framePosition - q_3d_ref, J_3d = frameRotation * J_frame(LOCAL).topRows<3>();
// they ride in the first three entries of diff / JJ (the angular half is zero, and so are its weights).
template <bool WITH_JACOBIAN>
__device__ __forceinline__ void task_evaluate(const double (&R)[9], V3 p, const double* __restrict__ ee,
                                              const double* __restrict__ ref, TaskEval& te, int kind) {
  // oMf = oMi[last joint] * frame placement
  double R6[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R6[k] = oct_bcast(R[k], NV - 1);
  const V3 p6 = V3{oct_bcast(p.x, NV - 1), oct_bcast(p.y, NV - 1), oct_bcast(p.z, NV - 1)};
  double Rf[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      Rf[3 * r + k] = fma(R6[3 * r + 2], ee[6 + k], fma(R6[3 * r + 1], ee[3 + k], R6[3 * r] * ee[k]));
  const V3 pf = V3{fma(R6[2], ee[11], fma(R6[1], ee[10], fma(R6[0], ee[9], p6.x))),
                   fma(R6[5], ee[11], fma(R6[4], ee[10], fma(R6[3], ee[9], p6.y))),
                   fma(R6[8], ee[11], fma(R6[7], ee[10], fma(R6[6], ee[9], p6.z)))};
  if (kind == 2) {
    te.diff[0] = pf.x - ref[9]; te.diff[1] = pf.y - ref[10]; te.diff[2] = pf.z - ref[11];
    te.diff[3] = 0.0; te.diff[4] = 0.0; te.diff[5] = 0.0;
    if (WITH_JACOBIAN) {
      const V3 Sw = V3{R[2], R[5], R[8]};
      const V3 Jl = mulT3(Rf, cross(p, Sw) + cross(Sw, pf));     // linear rows of getFrameJacobian(frame, LOCAL)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        te.JJ[r] = dot3r(Rf + 3 * r, Jl);                        // frameRotation * J_6d.topRows<3>()
        te.JJ[3 + r] = 0.0;
      }
    }
    return;
  }
  // diff_SE3 = SE3_ref^-1 * oMf
  double Rd[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      Rd[3 * r + k] = fma(ref[6 + r], Rf[6 + k], fma(ref[3 + r], Rf[3 + k], ref[r] * Rf[k]));
  const V3 pd = mulT3(ref, pf - V3{ref[9], ref[10], ref[11]});
  // pinocchio::log6
  double theta;
  const V3 w = log3_canon(Rd, theta);
  const double t2 = theta * theta;
  double st = 0.0, ct = 1.0;
  if (!(theta < TASK_TAYLOR)) canon_sincos(theta, &st, &ct);
  {
    double alpha, beta;
    if (theta < TASK_TAYLOR) {
      alpha = (1.0 - t2 / 12.0) - (t2 * t2) / 720.0;
      beta = 1.0 / 12.0 + t2 / 720.0;
    } else {
      alpha = (theta * st) / (2.0 * (1.0 - ct));
      beta = 1.0 / t2 - st / ((2.0 * theta) * (1.0 - ct));
    }
    const V3 v = fmav(beta * dot(w, pd), w, fmav(-0.5, cross(w, pd), alpha * pd));
    te.diff[0] = v.x; te.diff[1] = v.y; te.diff[2] = v.z;
    te.diff[3] = w.x; te.diff[4] = w.y; te.diff[5] = w.z;
  }
  if (!WITH_JACOBIAN) return;
  // pinocchio::Jlog6 = [[A, B], [0, A]]
  double A[9], B[9];
  jlog3_canon(theta, w, A);
  {
    double beta, bdot;
    if (theta < TASK_TAYLOR) {
      beta = 1.0 / 12.0 + t2 / 720.0;
      bdot = 1.0 / 360.0;
    } else {
      const double tinv = 1.0 / theta, t2inv = tinv * tinv;
      const double inv_2_2ct = 1.0 / (2.0 * (1.0 - ct));
      beta = t2inv - (st * tinv) * inv_2_2ct;
      bdot = -2.0 * (t2inv * t2inv) + ((1.0 + st * tinv) * t2inv) * inv_2_2ct;
    }
    const double wTp = dot(w, pd);
    const V3 v3t = (bdot * wTp) * w - fma(t2, bdot, 2.0 * beta) * pd;
    const V3 bw = beta * w;
    const double v3v[3] = {v3t.x, v3t.y, v3t.z}, bwv[3] = {bw.x, bw.y, bw.z}, wv[3] = {w.x, w.y, w.z},
                 pv[3] = {pd.x, pd.y, pd.z};
    double C[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) C[3 * r + k] = fma(bwv[r], pv[k], v3v[r] * wv[k]);
    const double dg = wTp * beta;
    C[0] += dg; C[4] += dg; C[8] += dg;
    add_skew(0.5 * pd, C);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) B[3 * r + k] = fma(C[3 * r + 2], A[6 + k], fma(C[3 * r + 1], A[3 + k], C[3 * r] * A[k]));
  }
  // getFrameJacobian(frame, LOCAL), column of this lane's joint: [Rf^T (S_l + S_w x pf); Rf^T S_w]
  const V3 Sw = V3{R[2], R[5], R[8]};
  const V3 Sl = cross(p, Sw);
  const V3 Jl = mulT3(Rf, Sl + cross(Sw, pf));
  const V3 Ja = mulT3(Rf, Sw);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    te.JJ[r] = dot3r(A + 3 * r, Jl) + dot3r(B + 3 * r, Ja);
    te.JJ[3 + r] = dot3r(A + 3 * r, Ja);
  }
}

// gradient sum_k JJ[k][lane] (w_k diff_k) and Gauss-Newton rows h[r] = sum_k JJ[k][r] (w_k JJ[k][lane]), r = 0..6,
// for a weight vector w6 (already in the reference's internal order).  `tile` = this octet's exchange tile
// [8][PAIR_TILE] holding every lane's JJ column in its first 6 entries (written by task_share_columns).
__device__ __forceinline__ void task_share_columns(int lane, const TaskEval& te, double* __restrict__ tile) {
  double* mine = tile + lane * PAIR_TILE;
#pragma unroll
  for (int k = 0; k < 6; ++k) mine[k] = te.JJ[k];
  __syncwarp();
}
__device__ __forceinline__ void task_gradient_hessian(const TaskEval& te, const double* __restrict__ w6,
                                                      const double* __restrict__ tile, double& g, double (&h)[NV]) {
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) acc = fma(te.JJ[k], w6[k] * te.diff[k], acc);
  g = acc;
  double wj[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) wj[k] = w6[k] * te.JJ[k];
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const double* o = tile + r * PAIR_TILE;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) t = fma(o[k], wj[k], t);
    h[r] = t;
  }
}
// sum_k (w_k diff_k) diff_k
__device__ __forceinline__ double task_weighted_sqnorm(const TaskEval& te, const double* __restrict__ w6) {
  double l = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) l += (w6[k] * te.diff[k]) * te.diff[k];
  return l;
}

}  // namespace idocp_b200

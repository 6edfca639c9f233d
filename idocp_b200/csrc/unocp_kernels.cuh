// unocp_kernels.cuh -- the batched Newton step of idocp's UnOCPSolver (fixed base, no contacts).
//
//   k_linearize   <-> SplitUnOCP::linearizeOCP / computeKKTResidual + squaredNormKKTResidual
//                     (reference include/idocp/unocp/split_unocp.hxx:69-99,141-174), one octet per
//                     (instance, stage)
//   k_riccati     <-> UnRiccatiRecursion backward/forward (src/unocp/unriccati_recursion.cpp:39-65,
//                     unocp/split_unriccati_factorizer.hxx:30-68, backward_unriccati_recursion_
//                     factorizer.hxx:29-89) + computeCondensedDirection + fraction-to-boundary
//                     (unocp_solver.cpp:96-115), one octet per instance, serial over the horizon
//   k_update      <-> updatePrimal / updateDual (unocp_solver.cpp:121-133)
//   k_kkt_sum     <-> UnOCPSolver::KKTError (unocp_solver.cpp:190-202)
//   k_init_constraints <-> UnOCPSolver::initConstraints (unocp_solver.cpp:59-70)
//
// Lane l < 7 of an octet owns joint l (all cost / constraint / state-equation algebra is lane
// local) and column l of every 7x7 block.
#pragma once
#include "chain_dynamics.cuh"

namespace idocp_b200 {

constexpr int OCTETS_PER_CTA = 16;  // 128 threads
constexpr int CTA_THREADS = OCTETS_PER_CTA * OCT;

// ---------------------------------------------------------------------------------------------
// constraints: primal-dual interior point rows of the six joint-limit components.
// Per lane: one row of each component (its joint).  comp: 0 pos-lo 1 pos-up 2 vel-lo 3 vel-up
// 4 trq-lo 5 trq-up.  (constraints/pdipm.hxx, src/constraints/joint_*_limit.cpp)
// ---------------------------------------------------------------------------------------------
struct LaneLimits {
  double qmin, qmax, vmax, umax;
};

__device__ __forceinline__ double con_residual(int comp, const LaneLimits& L, double q, double v, double u,
                                               double slack) {
  switch (comp) {
    case 0: return L.qmin - q + slack;
    case 1: return q - L.qmax + slack;
    case 2: return (-L.vmax) - v + slack;
    case 3: return v - L.vmax + slack;
    case 4: return (-L.umax) - u + slack;
    default: return u - L.umax + slack;
  }
}
__device__ __forceinline__ double con_margin(int comp, const LaneLimits& L, double q, double v, double u) {
  switch (comp) {
    case 0: return q - L.qmin;
    case 1: return L.qmax - q;
    case 2: return v - (-L.vmax);
    case 3: return L.vmax - v;
    case 4: return u - (-L.umax);
    default: return L.umax - u;
  }
}
__device__ __forceinline__ bool comp_active(int comp, int time_stage) {
  return comp < 2 ? pos_active(time_stage) : (comp < 4 ? vel_active(time_stage) : true);
}
// pdipm::FractionToBoundary (pdipm.hxx:52-73) for one row
__device__ __forceinline__ double fraction_row(double rate, double x, double dx, double cur) {
  const double f = -rate * (x / dx);
  if (f > 0.0 && f < 1.0 && f < cur) return f;
  return cur;
}

__device__ __forceinline__ LaneLimits load_limits(const DevProblem& P, int lane) {
  return LaneLimits{P.q_min[lane], P.q_max[lane], P.v_max[lane], P.u_max[lane]};
}

// ---------------------------------------------------------------------------------------------
// k_init_constraints: slack = margin (pushed above the barrier), dual = barrier / slack
// (pdipm::SetSlackAndDualPositive, pdipm.hxx:13-23)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS) k_init_constraints(const DevProblem* __restrict__ Pp, Layout L,
                                                                  int stage_offset) {
  const DevProblem& P = *Pp;
  const int lane = lane_in_octet();
  const long task = static_cast<long>(blockIdx.x) * OCTETS_PER_CTA + (threadIdx.x >> 3);
  const long ntask = static_cast<long>(L.N) * L.Bp;
  if (task >= ntask) return;
  const int i = static_cast<int>(task / L.Bp);
  const int b = static_cast<int>(task % L.Bp);
  const int ns = L.N + 1;
  const LaneLimits lim = load_limits(P, lane);
  const double q = L.sol[slot_index(S_Q, ns, i, L.Bp, b, lane)];
  const double v = L.sol[slot_index(S_V, ns, i, L.Bp, b, lane)];
  const double u = L.sol[slot_index(S_U, ns, i, L.Bp, b, lane)];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double sl = 0.0, du = 0.0;
    if (comp_active(c, i + stage_offset) && lane < NV) {
      sl = con_margin(c, lim, q, v, u);
      int guard = 0;
      while (sl < P.barrier && guard < (1 << 20)) { sl += P.barrier; ++guard; }
      du = P.barrier / sl;
    }
    L.slack[slot_index(c, L.N, i, L.Bp, b, lane)] = sl;
    L.dual[slot_index(c, L.N, i, L.Bp, b, lane)] = du;
  }
}

// set one solution field of every stage from value[b][7] (broadcast: value[7])
__global__ void k_set_solution(Layout L, int field, const double* __restrict__ value, int broadcast, int nstages) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = static_cast<long>(nstages) * L.Bp * OCT;
  if (idx >= total) return;
  const int lane = static_cast<int>(idx & 7);
  const long t = idx >> 3;
  const int b = static_cast<int>(t % L.Bp);
  const int i = static_cast<int>(t / L.Bp);
  double val = 0.0;
  if (lane < NV && b < L.B) val = broadcast ? value[lane] : value[static_cast<size_t>(b) * NV + lane];
  L.sol[slot_index(field, L.N + 1, i, L.Bp, b, lane)] = val;
}

// ---------------------------------------------------------------------------------------------
// k_linearize
// ---------------------------------------------------------------------------------------------
// RESIDUAL_ONLY = true : computeKKTResidual + squaredNormKKTResidual (writes kkt_stage only)
// RESIDUAL_ONLY = false: linearizeOCP (writes condensed KKT blocks, residual, expansion data)
template <bool RESIDUAL_ONLY>
__global__ void __launch_bounds__(CTA_THREADS) k_linearize(const DevProblem* __restrict__ Pp, Layout L) {
  IDOCP_DYN_SMEM(double, smem);
  const DevProblem& P = *Pp;
  const int lane = lane_in_octet();
  const int oct = threadIdx.x >> 3;
  double* tile = smem + oct * (OCT * PAIR_TILE);
  const int nstage_tasks = RESIDUAL_ONLY ? L.N + 1 : L.N;
  const long ntask = static_cast<long>(nstage_tasks) * L.Bp;
  long task = static_cast<long>(blockIdx.x) * OCTETS_PER_CTA + oct;
  if (task >= ntask) task = ntask - 1;  // tail octets redo the last task (keeps the warp convergent); stores are idempotent
  const int i = static_cast<int>(task / L.Bp);
  const int b = static_cast<int>(task % L.Bp);
  const int ns = L.N + 1;
  const int Bp = L.Bp;
  const double dt = P.dt;
  const bool act = lane < NV;

  const double q = L.sol[slot_index(S_Q, ns, i, Bp, b, lane)];
  const double v = L.sol[slot_index(S_V, ns, i, Bp, b, lane)];
  const double lmd = L.sol[slot_index(S_LMD, ns, i, Bp, b, lane)];
  const double gmm = L.sol[slot_index(S_GMM, ns, i, Bp, b, lane)];

  if (RESIDUAL_ONLY && i == L.N) {
    // TerminalOCP::computeKKTResidual + squaredNormKKTResidual (ocp/terminal_ocp.hxx:120-144)
    double lq = 0.0, lv = 0.0;
    lq += P.qf_weight[lane] * (q - P.q_ref[lane]);
    lv += P.vf_weight[lane] * (v - P.v_ref[lane]);
    lq -= lmd;
    lv -= gmm;
    if (!act) { lq = 0.0; lv = 0.0; }
    const double e = oct_sum_ordered(lq * lq) + oct_sum_ordered(lv * lv);
    if (lane == 0) L.kkt_stage[static_cast<size_t>(i) * Bp + b] = e;
    return;
  }

  const double a = L.sol[slot_index(S_A, ns, i, Bp, b, lane)];
  const double u = L.sol[slot_index(S_U, ns, i, Bp, b, lane)];
  const double beta = L.sol[slot_index(S_BETA, ns, i, Bp, b, lane)];
  const double qn = L.sol[slot_index(S_Q, ns, i + 1, Bp, b, lane)];
  const double vn = L.sol[slot_index(S_V, ns, i + 1, Bp, b, lane)];
  const double lmdn = L.sol[slot_index(S_LMD, ns, i + 1, Bp, b, lane)];
  const double gmmn = L.sol[slot_index(S_GMM, ns, i + 1, Bp, b, lane)];
  double slack[NC], dual[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    slack[c] = L.slack[slot_index(c, L.N, i, Bp, b, lane)];
    dual[c] = L.dual[slot_index(c, L.N, i, Bp, b, lane)];
  }
  const LaneLimits lim = load_limits(P, lane);

  // ---- inverse dynamics and its derivatives (UnconstrainedDynamics::linearizeInverseDynamics) ----
  JointDyn J;
  chain_world_sweep(lane, q, v, a, P.model + lane * MODEL_STRIDE, P.gravity, J);
  double dqc[NV], dvc[NV], Mc[NV];
  chain_pair_phase(lane, J, tile, dqc, dvc, Mc);
  const double ID = J.tau - u;

  // ---- gradient of the Lagrangian ----
  double lq = 0.0, lv = 0.0, la = 0.0, lu = 0.0;
  lq += dt * P.q_weight[lane] * (q - P.q_ref[lane]);
  lv += dt * P.v_weight[lane] * (v - P.v_ref[lane]);
  la += dt * P.a_weight[lane] * a;
  lu += dt * P.u_weight[lane] * (u - P.u_ref[lane]);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (!comp_active(c, i)) continue;
    const double g = dt * dual[c];
    const double sg = (c & 1) ? g : -g;
    if (c < 2) lq += sg; else if (c < 4) lv += sg; else lu += sg;
  }
  const double Fq = fma(dt, v, q - qn);
  const double Fv = fma(dt, a, v) - vn;
  lq += lmdn - lmd;
  lv += fma(dt, lmdn, gmmn) - gmm;
  la = fma(dt, gmmn, la);
  {
    double tq = 0.0, tv = 0.0, ta = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double bk = oct_bcast(beta, k);
      tq = fma(dqc[k], bk, tq);
      tv = fma(dvc[k], bk, tv);
      ta = fma(Mc[k], bk, ta);
    }
    lq = fma(dt, tq, lq);
    lv = fma(dt, tv, lv);
    la = fma(dt, ta, la);
    lu = fma(-dt, beta, lu);
  }

  if (RESIDUAL_ONLY) {
    // SplitUnOCP::squaredNormKKTResidual (split_unocp.hxx:164-174), canonical order
    double e = 0.0;
    const double z = act ? 1.0 : 0.0;
    e += oct_sum_ordered(z * (lq * lq)) + oct_sum_ordered(z * (lv * lv));
    e += oct_sum_ordered(z * (la * la));
    e += oct_sum_ordered(z * (lu * lu));
    e += oct_sum_ordered(z * (Fq * Fq)) + oct_sum_ordered(z * (Fv * Fv));
    e += dt * dt * oct_sum_ordered(z * (ID * ID));
    double c2 = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (!comp_active(c, i)) continue;
      const double r = con_residual(c, lim, q, v, u, slack[c]);
      const double dl = slack[c] * dual[c] - P.barrier;
      c2 += oct_sum_ordered(z * (r * r)) + oct_sum_ordered(z * (dl * dl));
    }
    e += dt * dt * c2;
    if (lane == 0) L.kkt_stage[static_cast<size_t>(i) * Bp + b] = e;
    return;
  }

  // ---- Hessian diagonals + constraint condensing (steps 5-6) ----
  double Qqq_d = dt * P.q_weight[lane];
  double Qvv_d = dt * P.v_weight[lane];
  const double Qaa_d = dt * P.a_weight[lane];
  double Quu_d = dt * P.u_weight[lane];
  if (act) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (!comp_active(c, i)) continue;
      const double r = con_residual(c, lim, q, v, u, slack[c]);
      const double dl = slack[c] * dual[c] - P.barrier;
      const double h = dt * dual[c] / slack[c];
      const double g = dt * (dual[c] * r - dl) / slack[c];
      const double sg = (c & 1) ? g : -g;
      if (c < 2) { Qqq_d += h; lq += sg; }
      else if (c < 4) { Qvv_d += h; lv += sg; }
      else { Quu_d += h; lu += sg; }
    }
  } else {
    Quu_d = 0.0;
  }
  // ---- eliminate u (step 7, unconstrained_dynamics.hxx:68-94) ----
  const double lu_c = act ? fma(Quu_d, ID, lu) : 0.0;
  double ulq = lq, ulv = lv, ula = la;
  {
    double tq = 0.0, tv = 0.0, ta = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double lk = oct_bcast(lu_c, k);
      tq = fma(dqc[k], lk, tq);
      tv = fma(dvc[k], lk, tv);
      ta = fma(Mc[k], lk, ta);
    }
    ulq += tq; ulv += tv; ula += ta;
  }
  const size_t Ns = L.N;
  L.kktR[slot_index(R_FQ, Ns, i, Bp, b, lane)] = Fq;
  L.kktR[slot_index(R_FV, Ns, i, Bp, b, lane)] = Fv;
  L.kktR[slot_index(R_LA, Ns, i, Bp, b, lane)] = ula;
  L.kktR[slot_index(R_LQ, Ns, i, Bp, b, lane)] = ulq;
  L.kktR[slot_index(R_LV, Ns, i, Bp, b, lane)] = ulv;
  L.expd[slot_index(E_ID, Ns, i, Bp, b, lane)] = ID;
  L.expd[slot_index(E_LU, Ns, i, Bp, b, lane)] = lu;
  L.expd[slot_index(E_QUU, Ns, i, Bp, b, lane)] = Quu_d;

  // exchange the columns of dID/dq, dID/dv, M through the tile: tile[lane][0..20]
  double* mine = tile + lane * PAIR_TILE;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    mine[k] = act ? dqc[k] : 0.0;
    mine[NV + k] = act ? dvc[k] : 0.0;
    mine[2 * NV + k] = act ? Mc[k] : 0.0;
  }
  __syncwarp();
  // expansion data in ROW layout (lane r holds row r): row r, col c = tile[c][r]
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const double* o = tile + c * PAIR_TILE;
    L.expd[slot_index(E_DQ + c, Ns, i, Bp, b, lane)] = o[lane < NV ? lane : 0];
    L.expd[slot_index(E_DV + c, Ns, i, Bp, b, lane)] = o[NV + (lane < NV ? lane : 0)];
    L.expd[slot_index(E_M + c, Ns, i, Bp, b, lane)] = Mc[c];
  }
  // own columns scaled by diag(Quu):  D*[k] = Quu_k * d*[k][c]
  double Dq[NV], Dv[NV], Da[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double qk = oct_bcast(Quu_d, k);
    Dq[k] = qk * dqc[k];
    Dv[k] = qk * dvc[k];
    Da[k] = qk * Mc[k];
  }
  // Q_xy[r][c] = sum_k d_x[k][r] * (Quu_k d_y[k][c])   (+ un-condensed diagonals)
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const double* o = tile + r * PAIR_TILE;
    double qq = 0.0, qv = 0.0, vv = 0.0, aq = 0.0, av = 0.0, aa = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double xq = o[k], xv = o[NV + k], xa = o[2 * NV + k];
      qq = fma(xq, Dq[k], qq);
      qv = fma(xq, Dv[k], qv);
      vv = fma(xv, Dv[k], vv);
      aq = fma(xa, Dq[k], aq);
      av = fma(xa, Dv[k], av);
      aa = fma(xa, Da[k], aa);
    }
    const bool diag = (r == lane);
    L.kktQ[slot_index(K_QQ * NV + r, Ns, i, Bp, b, lane)] = qq + (diag ? Qqq_d : 0.0);
    L.kktQ[slot_index(K_QV * NV + r, Ns, i, Bp, b, lane)] = qv;
    L.kktQ[slot_index(K_VV * NV + r, Ns, i, Bp, b, lane)] = vv + (diag ? Qvv_d : 0.0);
    L.kktQ[slot_index(K_AQ * NV + r, Ns, i, Bp, b, lane)] = aq;
    L.kktQ[slot_index(K_AV * NV + r, Ns, i, Bp, b, lane)] = av;
    L.kktQ[slot_index(K_AA * NV + r, Ns, i, Bp, b, lane)] = aa + (diag ? Qaa_d : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------
// k_riccati: backward Riccati recursion, forward recursion, costate / condensed directions and
// fraction-to-boundary step sizes.  One octet per instance; lane c owns column c.
// ---------------------------------------------------------------------------------------------
constexpr int RIC_TILE = 17;               // odd stride (doubles) -> conflict-free transposed reads
constexpr int RIC_SMEM_PER_OCT = 2 * OCT * RIC_TILE;

__global__ void __launch_bounds__(CTA_THREADS) k_riccati(const DevProblem* __restrict__ Pp, Layout L,
                                                         const double* __restrict__ q0,
                                                         const double* __restrict__ v0) {
  IDOCP_DYN_SMEM(double, smem);
  const DevProblem& P = *Pp;
  const int lane = lane_in_octet();
  const int oct = threadIdx.x >> 3;
  double* tA = smem + oct * RIC_SMEM_PER_OCT;         // [8][RIC_TILE]
  double* tB = tA + OCT * RIC_TILE;                   // [8][RIC_TILE]
  int b = blockIdx.x * OCTETS_PER_CTA + oct;
  const bool valid = b < L.B;
  if (b >= L.Bp) b = L.Bp - 1;
  const int N = L.N, Bp = L.Bp, ns = L.N + 1;
  const double dt = P.dt;
  const double dt2 = dt * dt;
  const bool act = lane < NV;
  const int ln = act ? lane : 0;
  int chol_fail = 0;

  // ---- terminal stage: P_N = diag(qf, vf), s_N = -l_N (unriccati_recursion.cpp:39-47) ----
  double Pqq[NV], Pqv[NV], Pvq[NV], Pvv[NV], sq, sv;
  {
    const double q = L.sol[slot_index(S_Q, ns, N, Bp, b, lane)];
    const double v = L.sol[slot_index(S_V, ns, N, Bp, b, lane)];
    const double lmd = L.sol[slot_index(S_LMD, ns, N, Bp, b, lane)];
    const double gmm = L.sol[slot_index(S_GMM, ns, N, Bp, b, lane)];
    double lq = 0.0, lv = 0.0;
    lq += P.qf_weight[lane] * (q - P.q_ref[lane]);
    lv += P.vf_weight[lane] * (v - P.v_ref[lane]);
    lq -= lmd;
    lv -= gmm;
    sq = -lq;
    sv = -lv;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      Pqq[r] = (r == lane) ? P.qf_weight[lane] : 0.0;
      Pvv[r] = (r == lane) ? P.vf_weight[lane] : 0.0;
      Pqv[r] = 0.0;
      Pvq[r] = 0.0;
    }
  }

  // ---- backward recursion ----
  for (int i = N - 1; i >= 0; --i) {
    double Qaa[NV], Qaq[NV], Qav[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      Qaa[r] = L.kktQ[slot_index(K_AA * NV + r, N, i, Bp, b, lane)];
      Qaq[r] = L.kktQ[slot_index(K_AQ * NV + r, N, i, Bp, b, lane)];
      Qav[r] = L.kktQ[slot_index(K_AV * NV + r, N, i, Bp, b, lane)];
    }
    const double Fq = L.kktR[slot_index(R_FQ, N, i, Bp, b, lane)];
    const double Fv = L.kktR[slot_index(R_FV, N, i, Bp, b, lane)];
    double la = L.kktR[slot_index(R_LA, N, i, Bp, b, lane)];
    const double lq = L.kktR[slot_index(R_LQ, N, i, Bp, b, lane)];
    const double lv = L.kktR[slot_index(R_LV, N, i, Bp, b, lane)];

    // factorizeKKTMatrix, a-blocks (backward_unriccati_recursion_factorizer.hxx:44-53); the
    // x-blocks are folded into the P update below (same operation order per entry)
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      Qaq[r] = fma(dt, Pvq[r], Qaq[r]);
      Qav[r] = fma(dt2, Pvq[r], Qav[r]);
      Qav[r] = fma(dt, Pvv[r], Qav[r]);
      Qaa[r] = fma(dt2, Pvv[r], Qaa[r]);
    }
    // products of the OLD P with Fx (used by la and by the s recursion)
    double pqqF = 0.0, pvqF = 0.0, pqvF = 0.0, pvvF = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double fq = oct_bcast(Fq, k), fv = oct_bcast(Fv, k);
      pqqF = fma(Pqq[k], fq, pqqF);   // (Pqq Fq)_c   (row c = column c, symmetric)
      pvqF = fma(Pvq[k], fv, pvqF);   // (Pqv Fv)_c
      pqvF = fma(Pqv[k], fq, pqvF);   // (Pqv^T Fq)_c
      pvvF = fma(Pvv[k], fv, pvvF);   // (Pvv Fv)_c
    }
    la = fma(dt, pqvF, la);
    la = fma(dt, pvvF, la);
    la = fma(-dt, sv, la);
    // share Qaa (tile A: tA[col][row]) and la
#pragma unroll
    for (int r = 0; r < NV; ++r) tA[lane * RIC_TILE + r] = Qaa[r];
    tA[lane * RIC_TILE + NV] = la;
    __syncwarp();
    // Cholesky of Qaa (lower triangle), redundantly in every lane: Lc[k][i], i >= k
    double Lm[NV][NV];  // Lm[col][row]; only row >= col used
    double rdiag[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double x = tA[k * RIC_TILE + k];
#pragma unroll
      for (int j = 0; j < k; ++j) x = fma(-Lm[j][k], Lm[j][k], x);
      if (!(x > 0.0)) chol_fail = 1;
      x = sqrt(x);
      Lm[k][k] = x;
      rdiag[k] = 1.0 / x;
#pragma unroll
      for (int r = k + 1; r < NV; ++r) {
        double y = tA[k * RIC_TILE + r];
#pragma unroll
        for (int j = 0; j < k; ++j) y = fma(-Lm[j][r], Lm[j][k], y);
        Lm[k][r] = y * rdiag[k];
      }
    }
    // K = -Qaa^-1 [Qaq Qav] (own columns), k = -Qaa^-1 la (all lanes)
    double Kq[NV], Kv[NV], kk[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double y1 = Qaq[r], y2 = Qav[r], y3 = tA[r * RIC_TILE + NV];
#pragma unroll
      for (int j = 0; j < r; ++j) {
        y1 = fma(-Lm[j][r], Kq[j], y1); y2 = fma(-Lm[j][r], Kv[j], y2); y3 = fma(-Lm[j][r], kk[j], y3);
      }
      Kq[r] = y1 * rdiag[r]; Kv[r] = y2 * rdiag[r]; kk[r] = y3 * rdiag[r];
    }
#pragma unroll
    for (int r = NV - 1; r >= 0; --r) {
      double y1 = Kq[r], y2 = Kv[r], y3 = kk[r];
#pragma unroll
      for (int j = r + 1; j < NV; ++j) {
        y1 = fma(-Lm[r][j], Kq[j], y1); y2 = fma(-Lm[r][j], Kv[j], y2); y3 = fma(-Lm[r][j], kk[j], y3);
      }
      Kq[r] = y1 * rdiag[r]; Kv[r] = y2 * rdiag[r]; kk[r] = y3 * rdiag[r];
    }
#pragma unroll
    for (int r = 0; r < NV; ++r) { Kq[r] = -Kq[r]; Kv[r] = -Kv[r]; kk[r] = -kk[r]; }
    // GK = Qaa K (full Qaa, as the reference multiplies the full matrix)
    double GKq[NV], GKv[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double g = tA[k * RIC_TILE + r];
        t1 = fma(g, Kq[k], t1);
        t2 = fma(g, Kv[k], t2);
      }
      GKq[r] = t1; GKv[r] = t2;
    }
    // share K (tile B: tB[col][0..6] = Kq[:,col], tB[col][7..13] = Kv[:,col])
#pragma unroll
    for (int r = 0; r < NV; ++r) { tB[lane * RIC_TILE + r] = Kq[r]; tB[lane * RIC_TILE + NV + r] = Kv[r]; }
    __syncwarp();
    // next s (uses the OLD P): backward_unriccati_recursion_factorizer.hxx:78-88
    double nsq, nsv;
    {
      nsq = sq; nsq -= pqqF; nsq -= pvqF;
      nsv = fma(dt, nsq, sv); nsv -= pqvF; nsv -= pvvF;
      nsq -= lq; nsv -= lv;
      double t5 = 0.0, t6 = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) { t5 = fma(Qaq[k], kk[k], t5); t6 = fma(Qav[k], kk[k], t6); }
      nsq -= t5; nsv -= t6;
    }
    // F-blocks of factorizeKKTMatrix (:34-43) + P = Qxx - K^T (Qaa K) (:64-75), row by row
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double Qqq = L.kktQ[slot_index(K_QQ * NV + r, N, i, Bp, b, lane)];
      double Qqv = L.kktQ[slot_index(K_QV * NV + r, N, i, Bp, b, lane)];
      double Qvv = L.kktQ[slot_index(K_VV * NV + r, N, i, Bp, b, lane)];
      Qqq += Pqq[r];
      Qqv = fma(dt, Pqq[r], Qqv);
      Qqv += Pqv[r];
      Qvv = fma(dt2, Pqq[r], Qvv);
      Qvv = fma(dt, Pqv[r], Qvv);
      Qvv = fma(dt, Pvq[r], Qvv);
      Qvv += Pvv[r];
      double tqq = 0.0, tqv = 0.0, tvv = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double kq = tB[r * RIC_TILE + k], kv = tB[r * RIC_TILE + NV + k];
        tqq = fma(kq, GKq[k], tqq);
        tqv = fma(kq, GKv[k], tqv);
        tvv = fma(kv, GKv[k], tvv);
      }
      Pqq[r] = Qqq - tqq;
      Pqv[r] = Qqv - tqv;
      Pvv[r] = Qvv - tvv;
    }
    // K in row layout for the forward pass: lane r gets Kq[r][c] = tB[c][r]
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      L.ric[slot_index(RC_KQ + c, N, i, Bp, b, lane)] = tB[c * RIC_TILE + ln];
      L.ric[slot_index(RC_KV + c, N, i, Bp, b, lane)] = tB[c * RIC_TILE + NV + ln];
    }
    L.ric[slot_index(RC_K, N, i, Bp, b, lane)] = kk[ln];
    __syncwarp();
    // transposes through the tiles: Pvq = Pqv^T, symmetrise Pqq and Pvv
#pragma unroll
    for (int r = 0; r < NV; ++r) { tA[lane * RIC_TILE + r] = Pqv[r]; tB[lane * RIC_TILE + r] = Pqq[r]; }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      Pvq[r] = tA[r * RIC_TILE + ln];                       // Pqv[lane][r]
      Pqq[r] = 0.5 * (Pqq[r] + tB[r * RIC_TILE + ln]);      // (Pqq[r][c] + Pqq[c][r]) / 2
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NV; ++r) tA[lane * RIC_TILE + r] = Pvv[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NV; ++r) Pvv[r] = 0.5 * (Pvv[r] + tA[r * RIC_TILE + ln]);
    __syncwarp();
    sq = nsq; sv = nsv;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      L.ric[slot_index(RC_PQQ + r, N, i, Bp, b, lane)] = Pqq[r];
      L.ric[slot_index(RC_PQV + r, N, i, Bp, b, lane)] = Pqv[r];
      L.ric[slot_index(RC_PVQ + r, N, i, Bp, b, lane)] = Pvq[r];
      L.ric[slot_index(RC_PVV + r, N, i, Bp, b, lane)] = Pvv[r];
    }
    L.ric[slot_index(RC_SQ, N, i, Bp, b, lane)] = sq;
    L.ric[slot_index(RC_SV, N, i, Bp, b, lane)] = sv;
  }

  // ---- forward recursion + directions + step sizes ----
  const LaneLimits lim = load_limits(P, lane);
  double dq, dv;
  {
    const size_t bi = static_cast<size_t>(valid ? b : 0) * NV + ln;
    dq = q0[bi] - L.sol[slot_index(S_Q, ns, 0, Bp, b, lane)];
    dv = v0[bi] - L.sol[slot_index(S_V, ns, 0, Bp, b, lane)];
    if (!act) { dq = 0.0; dv = 0.0; }
  }
  double min_p = 1.0, min_d = 1.0;
  for (int i = 0; i < N; ++i) {
    double dqk[NV], dvk[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) { dqk[k] = oct_bcast(dq, k); dvk[k] = oct_bcast(dv, k); }
    // da = K dx + k  (row layout)
    double da;
    {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < NV; ++c) acc = fma(L.ric[slot_index(RC_KQ + c, N, i, Bp, b, lane)], dqk[c], acc);
#pragma unroll
      for (int c = 0; c < NV; ++c) acc = fma(L.ric[slot_index(RC_KV + c, N, i, Bp, b, lane)], dvk[c], acc);
      da = acc + L.ric[slot_index(RC_K, N, i, Bp, b, lane)];
    }
    // costate direction (split_unriccati_factorizer.hxx:60-68)
    double dlmd, dgmm;
    {
      double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        t1 = fma(L.ric[slot_index(RC_PQQ + k, N, i, Bp, b, lane)], dqk[k], t1);
        t2 = fma(L.ric[slot_index(RC_PVQ + k, N, i, Bp, b, lane)], dvk[k], t2);
        t3 = fma(L.ric[slot_index(RC_PQV + k, N, i, Bp, b, lane)], dqk[k], t3);
        t4 = fma(L.ric[slot_index(RC_PVV + k, N, i, Bp, b, lane)], dvk[k], t4);
      }
      dlmd = t1; dlmd += t2; dlmd -= L.ric[slot_index(RC_SQ, N, i, Bp, b, lane)];
      dgmm = t3; dgmm += t4; dgmm -= L.ric[slot_index(RC_SV, N, i, Bp, b, lane)];
    }
    // du = ID + dID/dq dq + dID/dv dv + M da ; dbeta = (lu + Quu du) / dt
    double du, dbeta;
    {
      double acc = L.expd[slot_index(E_ID, N, i, Bp, b, lane)];
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < NV; ++c) t = fma(L.expd[slot_index(E_DQ + c, N, i, Bp, b, lane)], dqk[c], t);
      acc += t;
      t = 0.0;
#pragma unroll
      for (int c = 0; c < NV; ++c) t = fma(L.expd[slot_index(E_DV + c, N, i, Bp, b, lane)], dvk[c], t);
      acc += t;
      t = 0.0;
#pragma unroll
      for (int c = 0; c < NV; ++c) t = fma(L.expd[slot_index(E_M + c, N, i, Bp, b, lane)], oct_bcast(da, c), t);
      acc += t;
      du = acc;
      dbeta = fma(L.expd[slot_index(E_QUU, N, i, Bp, b, lane)], du, L.expd[slot_index(E_LU, N, i, Bp, b, lane)]) / dt;
    }
    // slack / dual directions and fraction-to-boundary
    if (act) {
      const double q = L.sol[slot_index(S_Q, ns, i, Bp, b, lane)];
      const double v = L.sol[slot_index(S_V, ns, i, Bp, b, lane)];
      const double u = L.sol[slot_index(S_U, ns, i, Bp, b, lane)];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (!comp_active(c, i)) continue;
        const double sl = L.slack[slot_index(c, N, i, Bp, b, lane)];
        const double dl = L.dual[slot_index(c, N, i, Bp, b, lane)];
        const double r = con_residual(c, lim, q, v, u, sl);
        const double dty = sl * dl - P.barrier;
        const double dx = c < 2 ? dq : (c < 4 ? dv : du);
        const double dslack = ((c & 1) ? -dx : dx) - r;
        const double ddual = -fma(dl, dslack, dty) / sl;
        min_p = fraction_row(P.fraction_rate, sl, dslack, min_p);
        min_d = fraction_row(P.fraction_rate, dl, ddual, min_d);
      }
    }
    L.dir[slot_index(D_LMD, ns, i, Bp, b, lane)] = dlmd;
    L.dir[slot_index(D_GMM, ns, i, Bp, b, lane)] = dgmm;
    L.dir[slot_index(D_Q, ns, i, Bp, b, lane)] = dq;
    L.dir[slot_index(D_V, ns, i, Bp, b, lane)] = dv;
    L.dir[slot_index(D_A, ns, i, Bp, b, lane)] = da;
    L.dir[slot_index(D_U, ns, i, Bp, b, lane)] = du;
    L.dir[slot_index(D_BETA, ns, i, Bp, b, lane)] = dbeta;
    // forwardRiccatiRecursion (split_unriccati_factorizer.hxx:49-57)
    double ndq = L.kktR[slot_index(R_FQ, N, i, Bp, b, lane)] + dq;
    double ndv = L.kktR[slot_index(R_FV, N, i, Bp, b, lane)] + dv;
    ndq = fma(dt, dv, ndq);
    ndv = fma(dt, da, ndv);
    dq = act ? ndq : 0.0;
    dv = act ? ndv : 0.0;
  }
  // terminal costate direction: P_N = diag, s_N recomputed
  {
    const double q = L.sol[slot_index(S_Q, ns, N, Bp, b, lane)];
    const double v = L.sol[slot_index(S_V, ns, N, Bp, b, lane)];
    const double lmd = L.sol[slot_index(S_LMD, ns, N, Bp, b, lane)];
    const double gmm = L.sol[slot_index(S_GMM, ns, N, Bp, b, lane)];
    double lq = 0.0, lv = 0.0;
    lq += P.qf_weight[lane] * (q - P.q_ref[lane]);
    lv += P.vf_weight[lane] * (v - P.v_ref[lane]);
    lq -= lmd;
    lv -= gmm;
    double dlmd = P.qf_weight[lane] * dq; dlmd -= -lq;
    double dgmm = P.vf_weight[lane] * dv; dgmm -= -lv;
    L.dir[slot_index(D_LMD, ns, N, Bp, b, lane)] = dlmd;
    L.dir[slot_index(D_GMM, ns, N, Bp, b, lane)] = dgmm;
    L.dir[slot_index(D_Q, ns, N, Bp, b, lane)] = dq;
    L.dir[slot_index(D_V, ns, N, Bp, b, lane)] = dv;
  }
  min_p = oct_min(min_p);
  min_d = oct_min(min_d);
  const int any_fail = __shfl_xor_sync(FULL, chol_fail, 1, OCT) | chol_fail;
  if (lane == 0 && valid) {
    L.steps[b] = min_p;
    L.steps[Bp + b] = min_d;
    int st = any_fail ? 1 : 0;
    if (!(min_p == min_p) || !(min_d == min_d)) st |= 2;
    L.status[b] |= st;
  }
}

// ---------------------------------------------------------------------------------------------
// k_update: s += alpha_p d, slack += alpha_p dslack, dual += alpha_d ddual
// (unocp_solver.cpp:121-133; split_solution.hxx:215-239; constraints_impl.hxx:181-196).
// dslack / ddual are recomputed from (s, slack, dual, d) instead of being stored.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS) k_update(const DevProblem* __restrict__ Pp, Layout L,
                                                        const double* __restrict__ primal_override) {
  const DevProblem& P = *Pp;
  const int lane = lane_in_octet();
  const long task = static_cast<long>(blockIdx.x) * OCTETS_PER_CTA + (threadIdx.x >> 3);
  const long ntask = static_cast<long>(L.N + 1) * L.Bp;
  if (task >= ntask) return;
  const int i = static_cast<int>(task / L.Bp);
  const int b = static_cast<int>(task % L.Bp);
  const int N = L.N, ns = L.N + 1, Bp = L.Bp;
  if (lane >= NV) return;
  const double ap = primal_override ? primal_override[b] : L.steps[b];
  const double ad = L.steps[Bp + b];
  const size_t iq = slot_index(S_Q, ns, i, Bp, b, lane), iv = slot_index(S_V, ns, i, Bp, b, lane);
  const size_t il = slot_index(S_LMD, ns, i, Bp, b, lane), ig = slot_index(S_GMM, ns, i, Bp, b, lane);
  const double q = L.sol[iq], v = L.sol[iv];
  const double dq = L.dir[slot_index(D_Q, ns, i, Bp, b, lane)];
  const double dv = L.dir[slot_index(D_V, ns, i, Bp, b, lane)];
  L.sol[il] = fma(ap, L.dir[slot_index(D_LMD, ns, i, Bp, b, lane)], L.sol[il]);
  L.sol[ig] = fma(ap, L.dir[slot_index(D_GMM, ns, i, Bp, b, lane)], L.sol[ig]);
  L.sol[iq] = fma(ap, dq, q);
  L.sol[iv] = fma(ap, dv, v);
  if (i == N) return;
  const size_t ia = slot_index(S_A, ns, i, Bp, b, lane), iu = slot_index(S_U, ns, i, Bp, b, lane);
  const size_t ib = slot_index(S_BETA, ns, i, Bp, b, lane);
  const double u = L.sol[iu];
  const double du = L.dir[slot_index(D_U, ns, i, Bp, b, lane)];
  L.sol[ia] = fma(ap, L.dir[slot_index(D_A, ns, i, Bp, b, lane)], L.sol[ia]);
  L.sol[iu] = fma(ap, du, u);
  L.sol[ib] = fma(ap, L.dir[slot_index(D_BETA, ns, i, Bp, b, lane)], L.sol[ib]);
  const LaneLimits lim = load_limits(P, lane);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (!comp_active(c, i)) continue;
    const size_t is = slot_index(c, N, i, Bp, b, lane);
    const double sl = L.slack[is], dl = L.dual[is];
    const double r = con_residual(c, lim, q, v, u, sl);
    const double dty = sl * dl - P.barrier;
    const double dx = c < 2 ? dq : (c < 4 ? dv : du);
    const double dslack = ((c & 1) ? -dx : dx) - r;
    const double ddual = -fma(dl, dslack, dty) / sl;
    L.slack[is] = fma(ap, dslack, sl);
    L.dual[is] = fma(ad, ddual, dl);
  }
}

// KKTError = sqrt(sum over stages in ascending order) (unocp_solver.cpp:190-202)
__global__ void k_kkt_sum(Layout L, int nstages) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L.Bp) return;
  double e = 0.0;
  for (int i = 0; i < nstages; ++i) e += L.kkt_stage[static_cast<size_t>(i) * L.Bp + b];
  L.kkt_err[b] = sqrt(e);
}

}  // namespace idocp_b200

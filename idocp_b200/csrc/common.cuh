// common.cuh -- device-side problem description and HBM layout of a batch of OCP instances.
#pragma once
#include "octet.cuh"

#ifdef IDOCP_B200_EMU
#define IDOCP_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_block->smem.data())
#define IDOCP_SINCOS(x, s, c) sincos_emu((x), (s), (c))
#else
#define IDOCP_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#define IDOCP_SINCOS(x, s, c) sincos((x), (s), (c))
#endif

namespace idocp_b200 {

constexpr int NC = 6;         // constraint components: pos lo/up, vel lo/up, torque lo/up
constexpr int MODEL_STRIDE = 24;  // doubles per joint in the model table

// per-joint model table (one row per lane, row 7 = padding joint with zero mass):
//  [0..8] placement R (row-major)  [9..11] placement p  [12] mass  [13..15] com  [16..21] inertia
struct DevProblem {
  int N;
  double T, dt;
  double q_ref[8], v_ref[8], u_ref[8];
  double q_weight[8], v_weight[8], a_weight[8], u_weight[8], qf_weight[8], vf_weight[8];
  double q_min[8], q_max[8], v_max[8], u_max[8];
  double barrier, fraction_rate;
  double gravity;
  double model[8 * MODEL_STRIDE];
};

// solution fields (one slot each, per stage)
enum SolField { S_LMD = 0, S_GMM, S_Q, S_V, S_A, S_U, S_BETA, S_NUM };
// condensed KKT blocks written by linearize: 7 slots (rows) each, lane = column
enum KktBlock { K_AA = 0, K_AQ, K_AV, K_QQ, K_QV, K_VV, K_NUMBLK };
// condensed residual slots
enum KktRes { R_FQ = 0, R_FV, R_LA, R_LQ, R_LV, R_NUM };
// expansion data slots written by linearize for the direction expansion:
//  ID, lu (after constraint condensing), Quu diag, rows of dID/dq (7), rows of dID/dv (7), col of M (7)
enum ExpSlot { E_ID = 0, E_LU, E_QUU, E_DQ = 3, E_DV = 10, E_M = 17, E_NUM = 24 };
// Riccati data kept for the forward pass: rows of Kq (7), rows of Kv (7), k, cols of Pqq, Pqv, Pvq, Pvv, sq, sv
enum RicSlot { RC_KQ = 0, RC_KV = 7, RC_K = 14, RC_PQQ = 15, RC_PQV = 22, RC_PVQ = 29, RC_PVV = 36, RC_SQ = 43, RC_SV = 44, RC_NUM = 45 };
// direction slots
enum DirField { D_LMD = 0, D_GMM, D_Q, D_V, D_A, D_U, D_BETA, D_NUM };

// All arrays are [slot][stage][instance(padded to 4)][8] doubles.
struct Layout {
  int B;     // instances
  int Bp;    // padded to a multiple of 4 (one warp = 4 octets)
  int N;     // stages with controls; solution has N+1 stages
  double* sol;    // [S_NUM][N+1]
  double* slack;  // [NC][N]
  double* dual;   // [NC][N]
  double* kktQ;   // [K_NUMBLK*7][N]
  double* kktR;   // [R_NUM][N]
  double* expd;   // [E_NUM][N]
  double* ric;    // [RC_NUM][N]
  double* dir;    // [D_NUM][N+1]
  double* steps;  // [2][Bp] primal, dual
  double* kkt_stage;  // [N+1][Bp] squared KKT norms per stage
  double* kkt_err;    // [Bp]
  int* status;        // [Bp]
};

__device__ __forceinline__ size_t slot_index(int slot, int nstage, int stage, int Bp, int b, int lane) {
  return ((static_cast<size_t>(slot) * nstage + stage) * Bp + b) * OCT + lane;
}

// constraint activity by time stage (reference constraints/constraints_data.hpp:18-43)
__device__ __forceinline__ bool pos_active(int time_stage) { return time_stage >= 2; }
__device__ __forceinline__ bool vel_active(int time_stage) { return time_stage >= 1; }

}  // namespace idocp_b200

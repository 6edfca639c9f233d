// chain_dynamics.cuh -- lane-parallel inverse dynamics (RNEA) and its analytical derivatives
// for a serial chain of revolute-Z joints (iiwa14), one joint per lane of an octet.
//
// Replaces Robot::RNEA / Robot::RNEADerivatives, i.e. pinocchio::rnea and
// pinocchio::computeRNEADerivatives (reference include/idocp/robot/robot.hxx:444-500).
//
// Formulation (DESIGN.md "RNEA derivatives"): all spatial quantities are expressed in the WORLD
// frame, [linear; angular].  Then the recursions over the chain become prefix / suffix sums over
// the lanes of the octet, and everything else is lane-local 3-vector algebra:
//   world transforms  : inclusive prefix PRODUCT of the local joint transforms (3 shuffle rounds)
//   v_i, a_i          : prefix sums of S_i qd_i and S_i qdd_i + dS_i qd_i
//   composites        : suffix sums of  I_i=(m,mc,Ibar) [10],  D_i=(hl,ha,Sym) [12],  f_i [6]
// The 6x6 composite "doYcrb" matrix of Carpentier & Mansard has only 12 independent entries:
//   D m = (-2 hl x m_w ; Sym m_w - ha x m_w).
// Pairwise phase (lane c owns column c):
//   r <= c: dq[r][c] = S_r.G_c    dv[r][c] = S_r.H_c    M[r][c] = S_r.U_c
//   r >  c: dq[r][c] = U_r.B_c + Ww_r.dSw_c   dv[r][c] = Ww_r.Sw_c + 2 U_r.dS_c   M[r][c] = S_c.U_r
//
// Every expression below is part of the canonical arithmetic (octet.cuh): the oracle's
// rnea_derivatives_impl repeats it operation by operation.
#pragma once
#include "common.cuh"

namespace idocp_b200 {

struct JointDyn {
  V3 Sl, Sw, dSl, dSw, Bl, Bw;
  V3 Ul, Uw, Ww, Gl, Gw, Hl, Hw;
  double tau;
};

// World placement of this lane's joint frame: local transform placement * Rz(q), then the inclusive
// prefix product X_l <- X_0 X_1 ... X_l over the chain (Hillis-Steele tree order).  R row-major
// (joint -> world), p = origin of the joint frame.  `mdl` points at this lane's row of the model table.
__device__ __forceinline__ void chain_fk(int lane, double q, const double* __restrict__ mdl, double (&R)[9], V3& p) {
  double sn, cs;
  canon_sincos(q, &sn, &cs);
  p = v3(mdl[9], mdl[10], mdl[11]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double a0 = mdl[3 * r], a1 = mdl[3 * r + 1];
    R[3 * r + 0] = fma(sn, a1, cs * a0);
    R[3 * r + 1] = fma(cs, a1, -(sn * a0));
    R[3 * r + 2] = mdl[3 * r + 2];
  }
#pragma unroll
  for (int d = 1; d < OCT; d <<= 1) {
    double Rs[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rs[k] = oct_up(R[k], d);
    const V3 ps = oct_up(p, d);
    if (lane >= d) {
      p = v3(fma(Rs[2], p.z, fma(Rs[1], p.y, fma(Rs[0], p.x, ps.x))),
             fma(Rs[5], p.z, fma(Rs[4], p.y, fma(Rs[3], p.x, ps.y))),
             fma(Rs[8], p.z, fma(Rs[7], p.y, fma(Rs[6], p.x, ps.z))));
      double Rn[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          Rn[3 * r + k] = fma(Rs[3 * r + 2], R[6 + k], fma(Rs[3 * r + 1], R[3 + k], Rs[3 * r] * R[k]));
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = Rn[k];
    }
  }
}

// Forward + backward world-frame sweeps on top of the placements of chain_fk.
// Lane 7 (padding) must be called with q = qd = qdd = 0; its model row has zero mass.
__device__ __forceinline__ void chain_world_sweep_from_fk(int lane, const double (&R)[9], V3 p, double qd, double qdd,
                                                          const double* __restrict__ mdl, double gravity,
                                                          JointDyn& J) {
  const ScanFlags sf = scan_flags(lane);
  const V3 z = v3(R[2], R[5], R[8]);
  J.Sl = cross(p, z);
  J.Sw = z;
  // velocities
  const V3 vw = oct_prefix_sum(qd * J.Sw, sf);
  const V3 vl = oct_prefix_sum(qd * J.Sl, sf);
  J.dSl = cross(vw, J.Sl) + cross(vl, J.Sw);
  J.dSw = cross(vw, J.Sw);
  // accelerations (gravity enters as the base acceleration (0,0,+g))
  const V3 aw = oct_prefix_sum(fmav(qd, J.dSw, qdd * J.Sw), sf);
  V3 al = oct_prefix_sum(fmav(qd, J.dSl, qdd * J.Sl), sf);
  al.z += gravity;
  J.Bl = ((cross(aw, J.Sl) + cross(al, J.Sw)) + cross(vw, J.dSl)) + cross(vl, J.dSw);
  J.Bw = cross(aw, J.Sw) + cross(vw, J.dSw);
  // world inertia about the world origin
  const double m = mdl[12];
  const V3 cm = v3(mdl[13], mdl[14], mdl[15]);
  const V3 cw = v3(fma(R[2], cm.z, fma(R[1], cm.y, fma(R[0], cm.x, p.x))),
                   fma(R[5], cm.z, fma(R[4], cm.y, fma(R[3], cm.x, p.y))),
                   fma(R[8], cm.z, fma(R[7], cm.y, fma(R[6], cm.x, p.z))));
  V3 mc = m * cw;
  S3 Ib;
  {
    const double i0 = mdl[16], i1 = mdl[17], i2 = mdl[18], i3 = mdl[19], i4 = mdl[20], i5 = mdl[21];
    double RI[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
      RI[3 * r + 0] = fma(r2, i2, fma(r1, i1, r0 * i0));
      RI[3 * r + 1] = fma(r2, i4, fma(r1, i3, r0 * i1));
      RI[3 * r + 2] = fma(r2, i5, fma(r1, i4, r0 * i2));
    }
    const double cc = dot(cw, cw);
    Ib.xx = fma(m, cc - cw.x * cw.x, fma(RI[2], R[2], fma(RI[1], R[1], RI[0] * R[0])));
    Ib.xy = fma(m, -(cw.x * cw.y), fma(RI[2], R[5], fma(RI[1], R[4], RI[0] * R[3])));
    Ib.xz = fma(m, -(cw.x * cw.z), fma(RI[2], R[8], fma(RI[1], R[7], RI[0] * R[6])));
    Ib.yy = fma(m, cc - cw.y * cw.y, fma(RI[5], R[5], fma(RI[4], R[4], RI[3] * R[3])));
    Ib.yz = fma(m, -(cw.y * cw.z), fma(RI[5], R[8], fma(RI[4], R[7], RI[3] * R[6])));
    Ib.zz = fma(m, cc - cw.z * cw.z, fma(RI[8], R[8], fma(RI[7], R[7], RI[6] * R[6])));
  }
  V3 hl = fmav(m, vl, cross(vw, mc));
  V3 ha = cross(mc, vl) + mul(Ib, vw);
  V3 fl = fmav(m, al, cross(aw, mc)) + cross(vw, hl);
  V3 fa = ((cross(mc, al) + mul(Ib, aw)) + cross(vw, ha)) + cross(vl, hl);
  S3 Sym;
  {
    // wI[r][k] = (vw x Ib[:,k])_r ;  Sym = -(vl mc^T + mc vl^T) + 2 (mc.vl) 1 + wI + wI^T
    const V3 c0 = cross(vw, v3(Ib.xx, Ib.xy, Ib.xz));
    const V3 c1 = cross(vw, v3(Ib.xy, Ib.yy, Ib.yz));
    const V3 c2 = cross(vw, v3(Ib.xz, Ib.yz, Ib.zz));
    const double mcv = dot(mc, vl);
    Sym.xx = 2.0 * (c0.x + (mcv - vl.x * mc.x));
    Sym.xy = (c1.x + c0.y) - fma(vl.x, mc.y, mc.x * vl.y);
    Sym.xz = (c2.x + c0.z) - fma(vl.x, mc.z, mc.x * vl.z);
    Sym.yy = 2.0 * (c1.y + (mcv - vl.y * mc.y));
    Sym.yz = (c2.y + c1.z) - fma(vl.y, mc.z, mc.y * vl.z);
    Sym.zz = 2.0 * (c2.z + (mcv - vl.z * mc.z));
  }
  // composite (suffix) sums
  const double mC = oct_suffix_sum(m, sf);
  mc = oct_suffix_sum(mc, sf);
  Ib.xx = oct_suffix_sum(Ib.xx, sf); Ib.xy = oct_suffix_sum(Ib.xy, sf); Ib.xz = oct_suffix_sum(Ib.xz, sf);
  Ib.yy = oct_suffix_sum(Ib.yy, sf); Ib.yz = oct_suffix_sum(Ib.yz, sf); Ib.zz = oct_suffix_sum(Ib.zz, sf);
  hl = oct_suffix_sum(hl, sf);
  ha = oct_suffix_sum(ha, sf);
  Sym.xx = oct_suffix_sum(Sym.xx, sf); Sym.xy = oct_suffix_sum(Sym.xy, sf); Sym.xz = oct_suffix_sum(Sym.xz, sf);
  Sym.yy = oct_suffix_sum(Sym.yy, sf); Sym.yz = oct_suffix_sum(Sym.yz, sf); Sym.zz = oct_suffix_sum(Sym.zz, sf);
  fl = oct_suffix_sum(fl, sf);
  fa = oct_suffix_sum(fa, sf);
  // per-joint vectors
  J.tau = dot(J.Sl, fl) + dot(J.Sw, fa);
  J.Ul = fmav(mC, J.Sl, cross(J.Sw, mc));
  J.Uw = cross(mc, J.Sl) + mul(Ib, J.Sw);
  J.Ww = fmav(2.0, cross(hl, J.Sl), mul(Sym, J.Sw)) + cross(ha, J.Sw);
  J.Gl = fmav(-2.0, cross(hl, J.dSw), fmav(mC, J.Bl, cross(J.Sw, fl)) + cross(J.Bw, mc));
  J.Gw = ((((cross(J.Sw, fa) + cross(J.Sl, fl)) + cross(mc, J.Bl)) + mul(Ib, J.Bw)) + mul(Sym, J.dSw)) - cross(ha, J.dSw);
  J.Hl = 2.0 * (fmav(mC, J.dSl, cross(J.dSw, mc)) - cross(hl, J.Sw));
  J.Hw = fmav(2.0, cross(mc, J.dSl) + mul(Ib, J.dSw), mul(Sym, J.Sw) - cross(ha, J.Sw));
}

__device__ __forceinline__ void chain_world_sweep(int lane, double q, double qd, double qdd,
                                                  const double* __restrict__ mdl, double gravity, JointDyn& J) {
  double R[9];
  V3 p;
  chain_fk(lane, q, mdl, R, p);
  chain_world_sweep_from_fk(lane, R, p, qd, qdd, mdl, gravity, J);
}

// tau only (used by the line search): same world sweep without the derivative vectors
__device__ __forceinline__ double chain_rnea(int lane, double q, double qd, double qdd,
                                             const double* __restrict__ mdl, double gravity) {
  JointDyn J;
  chain_world_sweep(lane, q, qd, qdd, mdl, gravity, J);
  return J.tau;
}

constexpr int PAIR_TILE = 25;  // doubles per lane in the exchange tile (odd: conflict-free transposes)

// Pairwise phase: produces column `lane` of dtau/dq, dtau/dv and M = dtau/da (rows r = 0..6).
// `tile` = this octet's shared-memory tile [8][PAIR_TILE].
__device__ __forceinline__ void chain_pair_phase(int lane, const JointDyn& J, double* __restrict__ tile,
                                                 double (&dqc)[NV], double (&dvc)[NV], double (&Mc)[NV]) {
  double* mine = tile + lane * PAIR_TILE;
  mine[0] = J.Sl.x; mine[1] = J.Sl.y; mine[2] = J.Sl.z;
  mine[3] = J.Sw.x; mine[4] = J.Sw.y; mine[5] = J.Sw.z;
  mine[6] = J.Ul.x; mine[7] = J.Ul.y; mine[8] = J.Ul.z;
  mine[9] = J.Uw.x; mine[10] = J.Uw.y; mine[11] = J.Uw.z;
  mine[12] = J.Ww.x; mine[13] = J.Ww.y; mine[14] = J.Ww.z;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const double* o = tile + r * PAIR_TILE;
    const V3 Sl = v3(o[0], o[1], o[2]), Sw = v3(o[3], o[4], o[5]);
    const V3 Ul = v3(o[6], o[7], o[8]), Uw = v3(o[9], o[10], o[11]);
    const V3 Ww = v3(o[12], o[13], o[14]);
    const double up_q = dot(Sl, J.Gl) + dot(Sw, J.Gw);
    const double up_v = dot(Sl, J.Hl) + dot(Sw, J.Hw);
    const double up_m = dot(Sl, J.Ul) + dot(Sw, J.Uw);
    const double lo_q = (dot(Ul, J.Bl) + dot(Uw, J.Bw)) + dot(Ww, J.dSw);
    const double lo_v = fma(2.0, dot(Ul, J.dSl) + dot(Uw, J.dSw), dot(Ww, J.Sw));
    const double lo_m = dot(J.Sl, Ul) + dot(J.Sw, Uw);
    const bool upper = (r <= lane);
    dqc[r] = upper ? up_q : lo_q;
    dvc[r] = upper ? up_v : lo_v;
    Mc[r] = upper ? up_m : lo_m;
  }
  __syncwarp();
}

}  // namespace idocp_b200

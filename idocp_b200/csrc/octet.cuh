// octet.cuh -- "octet" execution model helpers (sm_100a).
//
// One OCTET = 8 consecutive lanes of a warp works on one (instance, stage) pair or one
// instance; lane l < 7 owns joint l / matrix column l of the 7-dof iiwa14, lane 7 is padding.
// A warp therefore carries 4 independent problems; all cross-lane traffic stays inside the
// octet (width-8 shuffles, per-octet shared-memory tiles).
//
// HBM layout ("slots"): every per-joint vector is one 64-byte slot [8 doubles]; arrays are
// [slot][instance][8] so the 4 octets of a warp (4 consecutive instances) touch 256 contiguous
// bytes per load/store.
#pragma once
#ifndef IDOCP_B200_EMU
#include <cuda_runtime.h>
#endif

namespace idocp_b200 {

constexpr int NV = 7;        // iiwa14 nq = nv = nu
constexpr int OCT = 8;       // lanes per octet
constexpr unsigned FULL = 0xffffffffu;

struct V3 {
  double x, y, z;
};

__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// symmetric 3x3 (xx,xy,xz,yy,yz,zz)
struct S3 {
  double xx, xy, xz, yy, yz, zz;
};
__device__ __forceinline__ V3 mul(const S3& A, V3 b) {
  return V3{A.xx * b.x + A.xy * b.y + A.xz * b.z, A.xy * b.x + A.yy * b.y + A.yz * b.z,
            A.xz * b.x + A.yz * b.y + A.zz * b.z};
}

__device__ __forceinline__ int lane_in_octet() { return threadIdx.x & 7; }

// width-8 shuffles on doubles / V3
__device__ __forceinline__ double oct_up(double x, int d) { return __shfl_up_sync(FULL, x, d, OCT); }
__device__ __forceinline__ double oct_down(double x, int d) { return __shfl_down_sync(FULL, x, d, OCT); }
__device__ __forceinline__ double oct_bcast(double x, int src) { return __shfl_sync(FULL, x, src, OCT); }
__device__ __forceinline__ V3 oct_up(V3 a, int d) { return V3{oct_up(a.x, d), oct_up(a.y, d), oct_up(a.z, d)}; }
__device__ __forceinline__ V3 oct_down(V3 a, int d) {
  return V3{oct_down(a.x, d), oct_down(a.y, d), oct_down(a.z, d)};
}

// inclusive prefix sum over the octet (lane order 0..7)
__device__ __forceinline__ double oct_prefix_sum(double x, int lane) {
#pragma unroll
  for (int d = 1; d < OCT; d <<= 1) {
    const double y = oct_up(x, d);
    if (lane >= d) x += y;
  }
  return x;
}
__device__ __forceinline__ V3 oct_prefix_sum(V3 a, int lane) {
  return V3{oct_prefix_sum(a.x, lane), oct_prefix_sum(a.y, lane), oct_prefix_sum(a.z, lane)};
}
// inclusive suffix sum: x_l <- sum_{k >= l} x_k   (lane 7 must hold 0)
__device__ __forceinline__ double oct_suffix_sum(double x, int lane) {
#pragma unroll
  for (int d = 1; d < OCT; d <<= 1) {
    const double y = oct_down(x, d);
    if (lane + d < OCT) x += y;
  }
  return x;
}
__device__ __forceinline__ V3 oct_suffix_sum(V3 a, int lane) {
  return V3{oct_suffix_sum(a.x, lane), oct_suffix_sum(a.y, lane), oct_suffix_sum(a.z, lane)};
}
__device__ __forceinline__ double oct_min(double x) {
#pragma unroll
  for (int d = 1; d < OCT; d <<= 1) x = fmin(x, __shfl_xor_sync(FULL, x, d, OCT));
  return x;
}
// sum over lanes 0..6 in ASCENDING lane order (canonical reduction order of the oracle);
// every lane of the octet receives the result.
__device__ __forceinline__ double oct_sum_ordered(double x) {
  double s = oct_bcast(x, 0);
#pragma unroll
  for (int l = 1; l < NV; ++l) s += oct_bcast(x, l);
  return s;
}

}  // namespace idocp_b200

// dmma_probe.cu -- is the FP64 tensor-core MMA of sm_100a (mma.sync.aligned.m8n8k4.row.col.f64) usable inside the
// CANONICAL arithmetic of this library?  The parity scheme needs every dot product to be one
// ascending-index chain of IEEE fma operations.  This probe compares D = A B + C from the tensor core, bit for bit, with
//   (a) acc = c; for k = 0..3: acc = fma(a_k, b_k, acc)           (ascending chain)
//   (b) the same chain in descending k
//   (c) c + (exactly rounded sum of the four products) -- what a fused 4-term dot-product unit would give
// on random operands with widely spread exponents (so that the candidates differ from each other in most cases), and
// times a dependent DMMA chain against the equivalent DFMA work.  Output: one JSON line (profiles/dmma_probe.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// one warp per tile: A 8x4 row-major, B 4x8 (element (k, n) at B[k*8+n]), C / D 8x8 row-major
// fragment layout of m8n8k4 f64: a = A[lane/4][lane%4], b = B[lane%4][lane/4], c/d = C[lane/4][2*(lane%4) + {0,1}]
__global__ void k_probe(const double* A, const double* B, const double* C, double* D, int ntiles) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ntiles) return;
  const double* a = A + warp * 32;
  const double* b = B + warp * 32;
  const double* c = C + warp * 64;
  const int g = lane >> 2, t = lane & 3;
  double d0, d1;
  dmma_m8n8k4(d0, d1, a[g * 4 + t], b[t * 8 + g], c[g * 8 + 2 * t], c[g * 8 + 2 * t + 1]);
  D[warp * 64 + g * 8 + 2 * t] = d0;
  D[warp * 64 + g * 8 + 2 * t + 1] = d1;
}

// throughput: ITER dependent DMMAs per warp, 4 independent accumulators
__global__ void k_dmma_rate(double* out, int iters) {
  double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma_m8n8k4(c[2 * j], c[2 * j + 1], a, b, c[2 * j], c[2 * j + 1]);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma_rate(double* out, int iters) {
  double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = fma(a, b, c[j]);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t next64() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return rng_state;
}
static double rnd_spread() {   // sign * mantissa in [1,2) * 2^e, e in [-20, 20]
  const double m = 1.0 + (next64() >> 11) * (1.0 / 9007199254740992.0);
  const int e = static_cast<int>(next64() % 41) - 20;
  const double s = (next64() & 1) ? -1.0 : 1.0;
  return s * m * __builtin_ldexp(1.0, e);
}

int main() {
  const int ntiles = 4096;
  std::vector<double> A(ntiles * 32), B(ntiles * 32), C(ntiles * 64), D(ntiles * 64);
  for (auto& x : A) x = rnd_spread();
  for (auto& x : B) x = rnd_spread();
  for (auto& x : C) x = rnd_spread();
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
  k_probe<<<(ntiles * 32 + 127) / 128, 128>>>(dA, dB, dC, dD, ntiles);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
  long asc = 0, desc = 0, fused = 0, none = 0, distinct = 0, total = 0;
  for (int w = 0; w < ntiles; ++w)
    for (int m = 0; m < 8; ++m)
      for (int n = 0; n < 8; ++n) {
        const double* a = &A[w * 32 + m * 4];
        const double c = C[w * 64 + m * 8 + n];
        double x = c, y = c;
        for (int k = 0; k < 4; ++k) x = __builtin_fma(a[k], B[w * 32 + k * 8 + n], x);
        for (int k = 3; k >= 0; --k) y = __builtin_fma(a[k], B[w * 32 + k * 8 + n], y);
        __float128 z = c;
        for (int k = 0; k < 4; ++k) z += static_cast<__float128>(a[k]) * static_cast<__float128>(B[w * 32 + k * 8 + n]);
        const double zf = static_cast<double>(z);
        const double d = D[w * 64 + m * 8 + n];
        ++total;
        if (x != y || x != zf) ++distinct;
        if (d == x) ++asc;
        if (d == y) ++desc;
        if (d == zf) ++fused;
        if (d != x && d != y && d != zf) ++none;
      }
  // rates
  double* dout;
  cudaMalloc(&dout, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  float ms_mma = 0, ms_fma = 0;
  k_dmma_rate<<<148 * 8, 256>>>(dout, 100);
  k_dfma_rate<<<148 * 8, 256>>>(dout, 100);
  cudaEventRecord(e0); k_dmma_rate<<<148 * 8, 256>>>(dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms_mma, e0, e1);
  cudaEventRecord(e0); k_dfma_rate<<<148 * 8, 256>>>(dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms_fma, e0, e1);
  const double warps = 148.0 * 8 * 8;
  const double tf_mma = warps * iters * 4 * (8 * 8 * 4 * 2.0) / (ms_mma * 1e-3) / 1e12;
  const double tf_fma = warps * 32 * iters * 8 * 2.0 / (ms_fma * 1e-3) / 1e12;
  printf("{\"elements\": %ld, \"elements_where_candidates_differ\": %ld, \"equals_ascending_fma_chain\": %ld, "
         "\"equals_descending_fma_chain\": %ld, \"equals_exact_sum_rounded_once\": %ld, \"equals_none\": %ld, "
         "\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}\n",
         total, distinct, asc, desc, fused, none, tf_mma, tf_fma);
  return 0;
}

// fp64_peak.cu -- FP64 roofline denominators of THIS GPU, measured (BASELINE.md section 1 asks for a DFMA microbenchmark):
//   dfma       : independent DFMA chains on the CUDA cores (what every kernel of this repo issues)
//   dmma_m8n8k4: mma.sync.aligned.m8n8k4 f64 on the tensor cores (the only FP64 tensor route; tcgen05 has no FP64 kind)
// "burst" = best of 10 launches of ~2 ms each, timed alone; "sustained" = the average over >= 4 s of back-to-back launches
// (what a kernel inside a long step sees once the power limiter has settled).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Run:   tools/fp64_peak > profiles/fp64_peak.json
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

constexpr int CHAINS = 8;      // independent accumulators per thread
constexpr int INNER = 64;      // unrolled FMAs per chain per outer iteration

__global__ void __launch_bounds__(256) k_dfma(double* out, int outer, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int o = 0; o < outer; ++o) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

constexpr int MMA_CHAINS = 4;
__global__ void __launch_bounds__(256) k_dmma(double* out, int outer, double a, double b) {
  double c0[MMA_CHAINS], c1[MMA_CHAINS];
#pragma unroll
  for (int c = 0; c < MMA_CHAINS; ++c) { c0[c] = threadIdx.x * 1e-9 + c; c1[c] = c; }
  for (int o = 0; o < outer; ++o) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
#pragma unroll
      for (int c = 0; c < MMA_CHAINS; ++c)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < MMA_CHAINS; ++c) s += c0[c] + c1[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Result { double burst_tflops, sustained_tflops, sustained_seconds; int launches; };

template <typename F>
Result measure(F launch, double flops_per_launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int i = 0; i < 10; ++i) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, static_cast<double>(ms));
  }
  // sustained: >= 4 s back to back
  const int n = std::max(10, static_cast<int>(4000.0 / best) + 1);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < n; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  Result r;
  r.burst_tflops = flops_per_launch / (best * 1e-3) / 1e12;
  r.sustained_tflops = flops_per_launch * n / (ms * 1e-3) / 1e12;
  r.sustained_seconds = ms * 1e-3;
  r.launches = n;
  return r;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256;   // 8 CTAs x 8 warps = 64 warps / SM: the FP64 pipe never starves
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  const int outer = 400;
  const double fl_dfma = 2.0 * CHAINS * INNER * static_cast<double>(outer) * blocks * threads;
  // one m8n8k4 per warp = 8*8*4 FMAs = 512 flop
  const double fl_dmma = 512.0 * MMA_CHAINS * INNER * static_cast<double>(outer) * blocks * (threads / 32);
  Result a = measure([&] { k_dfma<<<blocks, threads>>>(out, outer, 0.999999, 1e-7); }, fl_dfma);
  Result b = measure([&] { k_dmma<<<blocks, threads>>>(out, outer, 0.999999, 1e-7); }, fl_dmma);
  CK(cudaGetLastError());
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  std::printf("{\"gpu_name\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f,\n", prop.name, sms, clock_khz / 1e3);
  std::printf(" \"fp64_dfma_tflops\": %.3f, \"fp64_dfma_tflops_sustained\": %.3f, \"dfma_sustained_seconds\": %.2f,\n",
              a.burst_tflops, a.sustained_tflops, a.sustained_seconds);
  std::printf(" \"fp64_dfma_per_clk_per_sm_at_max_clock\": %.2f,\n", a.burst_tflops * 1e12 / 2.0 / sms / (clock_khz * 1e3));
  std::printf(" \"fp64_dmma_m8n8k4_tflops\": %.3f, \"fp64_dmma_m8n8k4_tflops_sustained\": %.3f, \"dmma_sustained_seconds\": %.2f,\n",
              b.burst_tflops, b.sustained_tflops, b.sustained_seconds);
  std::printf(" \"how\": \"tools/fp64_peak.cu: %d CTAs x %d threads, %d independent DFMA chains per thread (%d mma.sync chains per warp), "
              "burst = best of 10 launches timed alone with CUDA events, sustained = >= 4 s back to back\"}\n",
              blocks, threads, CHAINS, MMA_CHAINS);
  return 0;
}

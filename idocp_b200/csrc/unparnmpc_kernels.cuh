// unparnmpc_kernels.cuh -- batched UnParNMPCSolver (backward-Euler stages, per-stage KKT
// inversion, backward/forward correction).  (reference src/unocp/unparnmpc_solver.cpp,
// src/unocp/unbackward_correction.cpp)
#pragma once
#include "unocp_kernels.cuh"

namespace idocp_b200 {

struct ParNMPCLayout {
  double* aux = nullptr;   // aux_mat per stage
};

template <typename Alloc>
inline int parnmpc_alloc(ParNMPCLayout& PL, int N, int Bp, Alloc alloc) {
  (void)PL; (void)N; (void)Bp; (void)alloc;
  return 0;
}

}  // namespace idocp_b200

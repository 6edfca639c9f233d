/* canon_pivot.h -- TEST INFRASTRUCTURE (CPU oracle): the canonical Cholesky pivot of the library.
 *
 * Eigen::LLT takes sqrt(pivot) and divides the column by it (Eigen/src/Cholesky/LLT.h, unblocked kernel).  The
 * canonical arithmetic shared by this oracle and the CUDA kernels (idocp_b200/csrc/octet.cuh: canon_rsqrt) instead
 * forms r ~ 1 / sqrt(x) directly and multiplies: L_ik = A_ik r, L_kk = x r.  The seed is the single-precision
 * 1.0f / sqrtf((float)x) -- two correctly rounded IEEE operations, bit-identical on any IEEE machine -- refined by
 * two Newton steps written with explicit fma (gcc -ffp-contract=off keeps them as written).  r is within about one
 * ulp of 1 / sqrt(x), i.e. the deviation from Eigen is of the same order as the reciprocal-multiply form it replaces
 * and far inside the unpinned Eigen boundary. */
#ifndef IDOCP_ORACLE_CANON_PIVOT_H
#define IDOCP_ORACLE_CANON_PIVOT_H
#include <math.h>

static inline int canon_pivot_ok(double x) { return x > 1e-36 && x < 1e36; }

static inline double canon_rsqrt(double x) {
  const float xf = (float)x;
  const float yf = 1.0f / sqrtf(xf);
  double y = (double)yf;
  const double h = 0.5 * x;
  double e = fma(-(h * y), y, 0.5);
  y = fma(y, e, y);
  e = fma(-(h * y), y, 0.5);
  y = fma(y, e, y);
  return y;
}

#endif

// Fused covariance builders for gpitch's kernels (reference: gpitch/matern12_spectral_mixture.py:38-62,102-133;
// GPflow-0.5 Stationary/Matern32/Add semantics per SURVEY.md Appendix A.1-A.2).
#pragma once
#include "common.cuh"

namespace gpx {

enum : int { KIND_MERCER_M12 = 0, KIND_DIFF_M12 = 1, KIND_MATERN32 = 2, KIND_DIFF_M32 = 3 };   // 3 shares 1's code paths
enum : int { DIST_REFERENCE = 0, DIST_STABLE = 1 };

struct KernArgs {
  int kind, mode;
  const double* ptsA;  // [ceil(batch/divA), nA] row points; batch entry b reads row b / divA
  const double* ptsB;  // [ceil(batch/divB), nB] column points; batch entry b reads row b / divB
  int nA, nB, divA, divB;
  const double* hyp;   // [batch, P, 2 + 2Q] = (variance, lengthscale, energy[Q], frequency[Q]) per component
  int P, Q;
  const double* featA;  // [batch, P, KP, nA]  (mercer only; KP = 2Q rounded up to a multiple of 4)
  const double* featB;  // [batch, P, KP, nB]
  double* K;            // out [batch, nA, ldk]   (build);   in: Kbar (grad)
  long long sK;
  int ldk;
  double jitter;        // added where global row == col (only meaningful when ptsA == ptsB)
  int batch;
  double* dhyp;         // grad only: [batch, P, 2 + 2Q], accumulated with atomics (caller zeroes)
  int need_ef;          // grad only: also produce energy / frequency gradients
  // grad only, optional fused epilogue on the incoming adjoint (all nullptr = plain Kbar):
  //   Kbar_eff[b,m,n] = epi_alpha * epi_col[b,n] * Kbar[b,m,n] + epi_rowv[b,m] * epi_colv[b,n]
  const double* epi_col;   // [batch, nB]
  const double* epi_rowv;  // [batch, nA]
  const double* epi_colv;  // [batch, nB]
  double epi_alpha;
  double* dpts;        // grad only, optional: [batch, nA] gradient w.r.t. the row points, accumulated (caller zeroes)
};

int launch_features(const double* pts, int n, int div, const double* hyp, int P, int Q, double* feat, int batch,
                    cudaStream_t st);
int launch_kernel_build(const KernArgs& a, cudaStream_t st);
int launch_kernel_grad(const KernArgs& a, cudaStream_t st);
int launch_kernel_grad_points(const KernArgs& a, double* dpts, cudaStream_t st);   // a.K holds Kbar
// lag-histogram gradient for inducing points on the sample grid (grad_lag.cu); a.K holds Kbar
int launch_kernel_grad_lag(const KernArgs& a, const int* iz, const double* delta, double* work, int nlag, cudaStream_t st);
inline int feat_rows(int Q) { return (2 * Q + 3) / 4 * 4; }

}  // namespace gpx

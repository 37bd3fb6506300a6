// Warp-reduced epilogues of the variational-GP inner loop: predictive-marginal column statistics, row dots,
// the Gauss-Hermite variational-expectation kernel and the whitened KL.
#pragma once
#include "common.cuh"

namespace gpx {
// fmean[b,n] = sum_m A[b,m,n] mu[b,m];  fvar[b,n] = kdiag[b] - sum_m A^2 + sum_m LTA^2   (GPflow conditional())
int launch_cond_colstats(const double* A, const double* LTA, long long sA, int ld, const double* mu,
                         const double* kdiag, double* fmean, double* fvar, int M, int N, int batch, int mode,
                         cudaStream_t st);
// out = alpha * colscale[n] * T + rowvec[m] colvec[n]
int launch_scale_rank1(const double* T, long long sT, int ld, const double* cs, const double* rv, const double* cv,
                       double alpha, double* out, int M, int N, int batch, cudaStream_t st);
// out[b,m] = sum_n A[b,m,n] v[b,n]
int launch_rowdot(const double* A, long long sA, int ld, const double* v, long long sV, double* out, int M, int N,
                  int batch, cudaStream_t st);
// MpdLik.variational_expectations fwd+bwd (likelihoods.py:33-68,422-447).  Fmu/Fvar/dFmu/dFvar [W, 2P, N]
// (rows 0..P-1 activations g, P..2P-1 components f), Y [W,N], noise [W]; ve_sum/dnoise [W] accumulated (caller zeroes).
int launch_varexp(const double* Fmu, const double* Fvar, const double* Y, const double* noise, int P, int W, int N,
                  int nlin, double* ve_sum, double* dFmu, double* dFvar, double* dnoise, double* ve_pointwise,
                  cudaStream_t st);
// merged_mean / merged_variance on the device (window_overlap.py:19-59); win = hann(ws) or hann(ws)^2 from the host
int launch_gather_cols(const double* Kuf, long long sF, int ldf, const int* iz, int div, int M, const double* pad_diag,
                       double jitter, double* Kuu, int batch, cudaStream_t st);
int launch_scatter_add_cols(const double* Kuu_bar, const int* iz, int div, int M, double* Kuf_bar, long long sF, int ldf,
                            int batch, cudaStream_t st);
int launch_tril_unpack(const double* packed, double* dense, int M, int batch, cudaStream_t st);
int launch_tril_pack(const double* dense, double* packed, int M, int batch, cudaStream_t st);
int launch_overlap_add(const double* Y, const double* win, int nw, int ws, int n, double* out, cudaStream_t st);
// gauss_kl(q_mu, q_sqrt) with K=None (whitened): kl[b], dmu[b,M], dLq[b,M,M] (lower; upper zeroed).
int launch_gauss_kl_white(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* dmu,
                          double* dLq, cudaStream_t st, double* tril_out = nullptr);
}

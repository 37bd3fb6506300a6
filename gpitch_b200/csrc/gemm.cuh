// Generic batched fp64 GEMM on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA.8x8x4), cp.async multi-stage
// pipeline.  One kernel template serves every dense contraction of the variational-GP inner loop:
//   TRMM  A = L^-1 Kuf, LTA = Lq^T A, Kbar = L^-T Abar     (sgpr_ss.py:48; conditional() in pdgp.py:147-155)
//   SYRK  AAT = A A^T, S_D = A diag(vbar) A^T             (sgpr_ss.py:49)
//   the M x M products of the Cholesky / inverse / adjoint recursions.
// All matrices row-major, contiguous last dim, leading dimensions in elements.
#pragma once
#include "common.cuh"

namespace gpx {

enum : int {
  GEMM_TRANS_A = 1,      // A operand stored [K, M] (use A^T)
  GEMM_TRANS_B = 2,      // B operand stored [N, K] (use B^T)
  GEMM_A_LOWER = 4,      // op(A)[m,k] == 0 for k > m  -> skip k-tiles beyond the row tile
  GEMM_A_UPPER = 8,      // op(A)[m,k] == 0 for k < m
  GEMM_B_LOWER = 16,     // op(B)[k,n] == 0 for k < n
  GEMM_B_UPPER = 32,     // op(B)[k,n] == 0 for k > n
  GEMM_C_LOWER = 64,     // only tiles touching the lower triangle of C are computed
  GEMM_C_MIRROR = 128,   // with C_LOWER: also write C[n,m] = C[m,n] (symmetric result)
  GEMM_ZERO_UPPER = 256  // with C_LOWER: write exact zeros above the diagonal inside computed tiles
};

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  long long sA, sB, sC;  // batch strides (elements); 0 = shared
  int lda, ldb, ldc;
  int M, N, K, batch;
  int flags;
  double alpha;             // C = colscale[n]*(alpha*alpha_vec[b]*acc + gamma*gamma_vec[b]*Aux[m,n]) + rowvec[m]*colvec[n] + beta*C
  double beta;
  double gamma;
  const double* alpha_vec;  // [batch] or null
  const double* kweight;    // [batch, K] weights applied along k (A diag(w) B) or null
  long long sKw;
  const double* Aux;        // [batch, M, N] (ld = ldaux) or null
  long long sAux;
  int ldaux;
  const double* colscale;   // [batch, N] or null
  const double* rowvec;     // [batch, M] or null
  const double* colvec;     // [batch, N] or null
  long long sColscale, sRowvec, sColvec;
  const double* gamma_vec;  // [batch] or null: per-batch factor on gamma (C = ... + gamma * gamma_vec[b] * Aux)
};

int launch_gemm(const GemmArgs& a, cudaStream_t st);
// TMA + mbarrier kernel (gemm_tma.cu): GPX_OK / GPX_ERR_* if it ran, 1 if the operands do not qualify for TMA.
int launch_gemm_tma(const GemmArgs& a, cudaStream_t st);

}  // namespace gpx

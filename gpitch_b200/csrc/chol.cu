#include "chol.cuh"
#include "gemm.cuh"
#include <cstdlib>

namespace gpx {

constexpr int NB = 64;          // diagonal block
constexpr int SLD = NB + 1;     // smem leading dim (odd -> column walks are conflict free)

// One CTA per matrix: factor the jb x jb diagonal block at (j0, j0), write L back (upper part of the block zeroed)
// and its inverse into the matching diagonal block of Linv.  The block lives in REGISTERS: 16 x 16 threads, thread
// (ty, tx) owns the 4 x 4 cyclic sub-grid (ty + 16 a, tx + 16 b); per elimination step only the pivot column (row, for
// the inverse) goes through a double-buffered shared-memory line, so a step costs one barrier and 16 FMAs per thread
// (the shared-memory version this replaces took ~100 us per block -- the latency that dominates single-window
// evaluations; blocks smaller than NB are padded with the identity).
__global__ void __launch_bounds__(256) diag_block_kernel(double* __restrict__ A, long long sA, int lda,
                                                         double* __restrict__ Linv, long long sI, int ldi, int j0,
                                                         int jb, int* __restrict__ info) {
  extern __shared__ __align__(16) double dsm[];
  double* S = dsm;                     // [NB][SLD] factor, for the inverse phase
  double* line = dsm + NB * SLD;       // [2][NB] pivot column / row, double buffered
  double* rdiag = line + 2 * NB;       // [NB] reciprocal pivots
  __shared__ int fail;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double* Ab = A + (long long)b * sA + (long long)j0 * lda + j0;
  double s[4][4], x[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int i = ty + 16 * a, j = tx + 16 * c;
      s[a][c] = (i < jb && j <= i) ? Ab[(long long)i * lda + j] : ((i == j) ? 1.0 : 0.0);
      x[a][c] = (i == j) ? 1.0 : 0.0;
    }
  if (tid == 0) fail = 0;
  // ---- Cholesky, right-looking, one column per step
#pragma unroll
  for (int kb = 0; kb < 4; kb++) {
    for (int kk = 0; kk < 16; kk++) {
      const int k = 16 * kb + kk;
      double* col = line + (k & 1) * NB;
      if (tx == kk) {                  // owners of column k publish it (rows >= k matter)
#pragma unroll
        for (int a = 0; a < 4; a++) col[ty + 16 * a] = s[a][kb];
      }
      __syncthreads();
      const double d = col[k];
      if (!(d > 0.0) && tid == 0 && fail == 0 && k < jb) fail = j0 + k + 1;      // also catches NaN
      // 1 / sqrt(d) on the critical path of every elimination step: hardware seed + two Newton steps (<= 1 ulp-ish)
      // instead of libm sqrt followed by an IEEE division (~3x the dependent latency); the step's pivot is d * rd.
      double rd;
      {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        const double hd = 0.5 * d;
        y = fma(y, fma(-hd * y, y, 0.5), y);
        rd = fma(y, fma(-hd * y, y, 0.5), y);
        if (!(d > 0.0)) rd = 1.0 / sqrt(d);      // failed matrix: keep IEEE NaN / inf semantics
      }
      if (tid == k) rdiag[k] = rd;
      double li[4], lj[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { li[a] = col[ty + 16 * a] * rd; lj[a] = col[tx + 16 * a] * rd; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int i = ty + 16 * a, j = tx + 16 * c;
          if (j > k && j <= i) s[a][c] -= li[a] * lj[c];
        }
      if (tx == kk) {                  // final column k of L
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const int i = ty + 16 * a;
          if (i == k) s[a][kb] = d * rd;
          else if (i > k) s[a][kb] = li[a];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int i = ty + 16 * a, j = tx + 16 * c;
      S[i * SLD + j] = (j <= i) ? s[a][c] : 0.0;
    }
  __syncthreads();
  // ---- inverse of the factor, right-looking: row k of X is final once steps < k are applied
#pragma unroll
  for (int kb = 0; kb < 4; kb++) {
    for (int kk = 0; kk < 16; kk++) {
      const int k = 16 * kb + kk;
      double* row = line + (k & 1) * NB;
      if (ty == kk) {                  // owners of row k: X[k][j] = x / L[k][k]
        const double rk = rdiag[k];          // 1 / L[k][k] from the factorisation phase
#pragma unroll
        for (int c = 0; c < 4; c++) { x[kb][c] *= rk; row[tx + 16 * c] = x[kb][c]; }
      }
      __syncthreads();
      double lk[4], xr[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { lk[a] = S[(ty + 16 * a) * SLD + k]; xr[a] = row[tx + 16 * a]; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int i = ty + 16 * a, j = tx + 16 * c;
          if (i > k && j <= k) x[a][c] -= lk[a] * xr[c];
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int i = ty + 16 * a, j = tx + 16 * c;
      if (i < jb && j < jb) {
        Ab[(long long)i * lda + j] = (j <= i) ? s[a][c] : 0.0;
        if (Linv) Linv[(long long)b * sI + (long long)(j0 + i) * ldi + j0 + j] = (j <= i) ? x[a][c] : 0.0;
      }
    }
  __syncthreads();
  if (tid == 0 && fail != 0 && info[b] == 0) info[b] = fail;
}

// Diagonal-block inverse only stored in a scratch block when the caller does not want Linv: handled by passing a
// scratch Linv (the panel solve needs the block inverse either way).

// Strict upper triangle <- 0: one warp per row (rows strided over the grid), lanes over the columns right of the diagonal.
__global__ void __launch_bounds__(256) zero_upper_kernel(double* __restrict__ A, long long sA, int lda, int M) {
  double* Ab = A + (long long)blockIdx.y * sA;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + warp; i < M; i += gridDim.x * 8) {
    double* row = Ab + (long long)i * lda;
    for (int j = i + 1 + lane; j < M; j += 32) row[j] = 0.0;
  }
}

static GemmArgs base_args(int batch) {
  GemmArgs g = {};
  g.batch = batch;
  g.alpha = 1.0;
  g.beta = 0.0;
  g.gamma = 0.0;
  return g;
}

static int zero_upper(double* A, long long sA, int lda, int M, int batch, cudaStream_t st) {
  for (int b0 = 0; b0 < batch; b0 += 32768) {  // grid.y limit
    const int nb_ = batch - b0 < 32768 ? batch - b0 : 32768;
    dim3 gz((unsigned)((M + 7) / 8 < 64 ? (M + 7) / 8 : 64), nb_);
    zero_upper_kernel<<<gz, 256, 0, st>>>(A + (long long)b0 * sA, sA, lda, M);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}

// Few large matrices (configs[4]: one window, M = 2048): the left-looking panel update below runs (M - j) / 128 CTAs with
// a k-loop as long as the factorised part -- a handful of CTAs doing long serial chains.  Here every step exposes the
// whole trailing matrix instead (right-looking: A22 -= L21 L21^T, (M - j)^2 / 2 outputs, K = 64), and the inverse is
// assembled recursively: with X = L^-1 = [[X11, 0], [-X22 L21 X11, X22]] the diagonal-block inverses are doubled in size
// level by level, all block pairs of a level in one batched launch (batch stride = the pair pitch along the diagonal).
// T' = X11^T L21^T is parked in the mirror block above the diagonal (zeroed at the end), so no extra workspace is needed.
static int potrf_trinv_wide(double* A, long long sA, int lda, double* Linv, long long sI, int ldi, int* info, int M,
                            int batch, cudaStream_t st) {
  const size_t DIAG_SMEM = (NB * SLD + 3 * NB) * sizeof(double);
  int rc;
  for (int j0 = 0; j0 < M; j0 += NB) {
    const int jb = (M - j0 < NB) ? (M - j0) : NB;
    diag_block_kernel<<<batch, 256, DIAG_SMEM, st>>>(A, sA, lda, Linv, sI, ldi, j0, jb, info);
    GPX_CHECK_LAUNCH();
    const int rest = M - j0 - jb;
    if (rest <= 0) break;
    GemmArgs g = base_args(batch);         // panel: L21 = A21 Dinv^T (in place; each CTA owns full rows)
    g.A = A + (long long)(j0 + jb) * lda + j0; g.sA = sA; g.lda = lda;
    g.B = Linv + (long long)j0 * ldi + j0; g.sB = sI; g.ldb = ldi;
    g.C = A + (long long)(j0 + jb) * lda + j0; g.sC = sA; g.ldc = lda;
    g.M = rest; g.N = jb; g.K = jb;
    g.flags = GEMM_TRANS_B | GEMM_B_UPPER;
    if ((rc = launch_gemm(g, st)) != GPX_OK) return rc;
    GemmArgs u = base_args(batch);         // trailing update: A22 -= L21 L21^T (lower triangle only)
    u.A = A + (long long)(j0 + jb) * lda + j0; u.sA = sA; u.lda = lda;
    u.B = u.A; u.sB = sA; u.ldb = lda;
    u.C = A + (long long)(j0 + jb) * lda + j0 + jb; u.sC = sA; u.ldc = lda;
    u.M = rest; u.N = rest; u.K = jb;
    u.flags = GEMM_TRANS_B | GEMM_C_LOWER;
    u.alpha = -1.0; u.beta = 1.0;
    if ((rc = launch_gemm(u, st)) != GPX_OK) return rc;
  }
  // recursive inverse: diagonal blocks of size h are final; build the blocks of size 2h
  for (int h = NB; h < M; h *= 2) {
    const int pitch = 2 * h;
    const int full = M / pitch;                              // pairs whose right block is complete
    const int rag = (M - full * pitch > h) ? (M - full * pitch - h) : 0;   // rows of a ragged last right block
    for (int pass = 0; pass < 2; pass++) {
      const int npair = pass == 0 ? full : (rag ? 1 : 0);
      if (npair == 0) continue;
      const int hr = pass == 0 ? h : rag;
      const long long o = pass == 0 ? 0 : (long long)full * pitch;          // first row / column of the pair(s)
      for (int b = 0; b < batch; b++) {                      // (few matrices by construction)
        double* Lb = A + (long long)b * sA;
        double* Xb = Linv + (long long)b * sI;
        GemmArgs t = base_args(npair);       // T' = X11^T L21^T  -> mirror block [h x hr] at (o, o + h)
        t.A = Xb + o * ldi + o; t.sA = (long long)pitch * (ldi + 1); t.lda = ldi;
        t.B = Lb + (o + h) * lda + o; t.sB = (long long)pitch * (lda + 1); t.ldb = lda;
        t.C = Xb + o * ldi + o + h; t.sC = (long long)pitch * (ldi + 1); t.ldc = ldi;
        t.M = h; t.N = hr; t.K = h;
        t.flags = GEMM_TRANS_A | GEMM_TRANS_B | GEMM_A_UPPER;
        if ((rc = launch_gemm(t, st)) != GPX_OK) return rc;
        GemmArgs x = base_args(npair);       // X21 = -X22 T'^T   -> block [hr x h] at (o + h, o)
        x.A = Xb + (o + h) * ldi + o + h; x.sA = (long long)pitch * (ldi + 1); x.lda = ldi;
        x.B = Xb + o * ldi + o + h; x.sB = (long long)pitch * (ldi + 1); x.ldb = ldi;
        x.C = Xb + (o + h) * ldi + o; x.sC = (long long)pitch * (ldi + 1); x.ldc = ldi;
        x.M = hr; x.N = h; x.K = hr;
        x.flags = GEMM_TRANS_B | GEMM_A_LOWER;
        x.alpha = -1.0;
        if ((rc = launch_gemm(x, st)) != GPX_OK) return rc;
      }
    }
    // clear the parked T' blocks: the next level reads these diagonal blocks as triangular operands, and the MMA blocks
    // that straddle their diagonal do touch elements above it
    if ((rc = zero_upper(Linv, sI, ldi, M, batch, st)) != GPX_OK) return rc;
  }
  if (M <= NB && (rc = zero_upper(Linv, sI, ldi, M, batch, st)) != GPX_OK) return rc;
  return zero_upper(A, sA, lda, M, batch, st);
}

int potrf_trinv(double* A, long long sA, int lda, double* Linv, long long sI, int ldi, double* work, int* info,
                int M, int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  if (!A || !Linv || !info || !work) return GPX_ERR_ARG;
  cudaMemsetAsync(info, 0, sizeof(int) * (size_t)batch, st);
  const size_t DIAG_SMEM = (NB * SLD + 3 * NB) * sizeof(double);
  cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  static const int wide_on = getenv("GPX_POTRF_WIDE") ? atoi(getenv("GPX_POTRF_WIDE")) : 1;
  if (wide_on && M >= 512 && (long long)batch * ((M + 127) / 128) <= 96)
    return potrf_trinv_wide(A, sA, lda, Linv, sI, ldi, info, M, batch, st);
  for (int b0 = 0; b0 < batch; b0 += 32768) {  // grid.y limit for the helper kernels
    const int nb_ = batch - b0 < 32768 ? batch - b0 : 32768;
    dim3 gz((unsigned)((M + 7) / 8 < 64 ? (M + 7) / 8 : 64), nb_);
    zero_upper_kernel<<<gz, 256, 0, st>>>(Linv + (long long)b0 * sI, sI, ldi, M);
    ++g_launches;
  }
  // Linv diagonal / lower parts are fully overwritten below; its strict upper triangle was just zeroed.
  int rc;
  for (int j0 = 0; j0 < M; j0 += NB) {
    const int jb = (M - j0 < NB) ? (M - j0) : NB;
    if (j0 > 0) {
      // left-looking update of block column j:  A[j0:, j0:j0+jb] -= L[j0:, 0:j0] L[j0:j0+jb, 0:j0]^T
      GemmArgs g = base_args(batch);
      g.A = A + (long long)j0 * lda; g.sA = sA; g.lda = lda;
      g.B = A + (long long)j0 * lda; g.sB = sA; g.ldb = lda;
      g.C = A + (long long)j0 * lda + j0; g.sC = sA; g.ldc = lda;
      g.M = M - j0; g.N = jb; g.K = j0;
      g.flags = GEMM_TRANS_B;
      g.alpha = -1.0; g.beta = 1.0;
      if ((rc = launch_gemm(g, st)) != GPX_OK) return rc;
    }
    diag_block_kernel<<<batch, 256, DIAG_SMEM, st>>>(A, sA, lda, Linv, sI, ldi, j0, jb, info);
    GPX_CHECK_LAUNCH();
    if (j0 + jb < M) {
      // panel: A[j0+jb:, j0:j0+jb] <- A[j0+jb:, j0:j0+jb] Dinv^T.  In place: each CTA owns full rows (jb <= BN).
      GemmArgs g = base_args(batch);
      g.A = A + (long long)(j0 + jb) * lda + j0; g.sA = sA; g.lda = lda;
      g.B = Linv + (long long)j0 * ldi + j0; g.sB = sI; g.ldb = ldi;
      g.C = A + (long long)(j0 + jb) * lda + j0; g.sC = sA; g.ldc = lda;
      g.M = M - j0 - jb; g.N = jb; g.K = jb;
      g.flags = GEMM_TRANS_B | GEMM_B_UPPER;
      if ((rc = launch_gemm(g, st)) != GPX_OK) return rc;
    }
  }
  for (int b0 = 0; b0 < batch; b0 += 32768) {
    const int nb_ = batch - b0 < 32768 ? batch - b0 : 32768;
    dim3 gz((unsigned)((M + 7) / 8 < 64 ? (M + 7) / 8 : 64), nb_);
    zero_upper_kernel<<<gz, 256, 0, st>>>(A + (long long)b0 * sA, sA, lda, M);
    ++g_launches;
  }
  if (cudaGetLastError() != cudaSuccess) return GPX_ERR_LAUNCH;
  // off-diagonal blocks of the inverse, block row by block row:
  //   Linv[i, 0:i0] = -Dinv_i ( L[i, 0:i0] Linv[0:i0, 0:i0] )
  for (int i0 = NB; i0 < M; i0 += NB) {
    const int ib = (M - i0 < NB) ? (M - i0) : NB;
    GemmArgs g = base_args(batch);
    g.A = A + (long long)i0 * lda; g.sA = sA; g.lda = lda;
    g.B = Linv; g.sB = sI; g.ldb = ldi;
    g.C = work; g.sC = (long long)NB * M; g.ldc = M;
    g.M = ib; g.N = i0; g.K = i0;
    g.flags = GEMM_B_LOWER;
    if ((rc = launch_gemm(g, st)) != GPX_OK) return rc;
    GemmArgs h = base_args(batch);
    h.A = Linv + (long long)i0 * ldi + i0; h.sA = sI; h.lda = ldi;
    h.B = work; h.sB = (long long)NB * M; h.ldb = M;
    h.C = Linv + (long long)i0 * ldi; h.sC = sI; h.ldc = ldi;
    h.M = ib; h.N = i0; h.K = ib;
    h.flags = GEMM_A_LOWER;
    h.alpha = -1.0;
    if ((rc = launch_gemm(h, st)) != GPX_OK) return rc;
  }
  return GPX_OK;
}

}  // namespace gpx

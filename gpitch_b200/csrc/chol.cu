#include "chol.cuh"
#include "gemm.cuh"
#include <cstdlib>
#include <cstdint>

namespace gpx {

constexpr int NB = 64;          // diagonal block
constexpr int SLD = NB + 1;     // smem leading dim (odd -> column walks are conflict free)

constexpr int PB = 16;          // inner panel of the diagonal-block kernel
#ifndef GPX_DIAG_STAMP            // tools/diag_probe.cu defines it to record clock64() per phase
#define GPX_DIAG_STAMP(i)
#endif
constexpr int LBS = PB + 2;     // pivot line: 16 column entries + the next pivot (+ pad)
constexpr size_t DIAG_SMEM = (2 * NB * SLD + 2 * LBS + NB) * sizeof(double);

// 1 / sqrt(d) sits on the critical path of every elimination step: hardware seed (MUFU.RSQ64H, ~2^-22) + ONE third-order
// step  y (1 + e/2 + 3 e^2 / 8),  e = 1 - d y^2  (error ~ e^3: below 2^-60) -- four dependent FP64 operations instead of
// the six of two Newton steps, and ~3x shorter than libm sqrt followed by an IEEE division.  The step's pivot is d * rd.
__device__ __forceinline__ double rsqrt_fast(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#if defined(GPX_RSQRT_NEWTON)
  const double hd = 0.5 * d;
  y = fma(y, fma(-hd * y, y, 0.5), y);
  y = fma(y, fma(-hd * y, y, 0.5), y);
#else
  const double e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
#endif
  if (!(d >= 2.2250738585072014e-308)) y = 1.0 / sqrt(d);      // failed matrix (or a subnormal pivot, which the ftz seed flushes): IEEE semantics
  return y;
}

// One CTA per matrix: factor the jb x jb diagonal block at (j0, j0), write L back (upper part of the block zeroed)
// and its inverse into the matching diagonal block of Linv.  A 64-column elimination is a chain of 64 dependent
// (pivot -> rsqrt -> scale -> update) steps, so the kernel is built to make each step short and to keep everything else
// off the chain.  The block is factored in 16-wide panels.  ONE WARP factors the 16 x 16 diagonal sub-block with a row
// per lane in registers and uniform control flow: every lane reads the pivot d_k from a shared line, computes
// 1/sqrt(d_k) itself, scales its own entry l_rk and publishes it; lane k+1 also publishes the NEXT pivot
// d_{k+1} = a_{k+1,k+1} - l_{k+1,k}^2, which needs nothing from the other lanes -- so the chain per step is
// LDS -> rsqrt -> DMUL -> DFMA -> STS -> __syncwarp, and the 15 update FMAs of the step overlap the next rsqrt.  The
// panel below is a forward substitution with one thread per row, the trailing update 16-term dot products register-tiled
// over all 256 threads; the inverse is built after the factorisation, again by substitution (see there).  Panels past
// jb (a ragged last block) are skipped.  (Measured with tools/diag_probe.cu; the first version eliminated the whole 64-wide block column by column
// with a block barrier per step.)
__global__ void __launch_bounds__(256) diag_block_kernel(double* __restrict__ A, long long sA, int lda,
                                                         double* __restrict__ Linv, long long sI, int ldi, int j0,
                                                         int jb, int* __restrict__ info) {
  extern __shared__ __align__(16) double dsm[];
  double* S = dsm;                     // [NB][SLD] block -> L
  double* X = S + NB * SLD;            // [NB][SLD] inverse
  double* LB = X + NB * SLD;           // [2][LBS] pivot column + next pivot, double buffered
  double* rdg = LB + 2 * LBS;          // [NB] reciprocal pivots
  __shared__ int fail;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
  const int np = (jb + PB - 1) / PB;   // live panels; rows >= nr are not touched at all
  const int nr = np * PB;
  double* Ab = A + (long long)b * sA + (long long)j0 * lda + j0;
  // one warp per row, a lane per column pair: 16-byte global accesses when the row pitch and base allow it
  const bool vecA = ((lda & 1) == 0) && ((((uintptr_t)Ab) & 15) == 0);
#pragma unroll 4
  for (int i = warp; i < nr; i += 8) {
    const int j = 2 * lane;
    double v0 = 0.0, v1 = 0.0;
    if (i < jb && j <= i) {
      if (vecA && j + 1 < jb) {
        const double2 t = *reinterpret_cast<const double2*>(Ab + (long long)i * lda + j);
        v0 = t.x; v1 = t.y;
      } else {
        v0 = Ab[(long long)i * lda + j];
        if (j + 1 < jb) v1 = Ab[(long long)i * lda + j + 1];
      }
    }
    S[i * SLD + j] = (j <= i) ? ((i < jb) ? v0 : (i == j ? 1.0 : 0.0)) : 0.0;
    S[i * SLD + j + 1] = (j + 1 <= i) ? ((i < jb) ? v1 : (i == j + 1 ? 1.0 : 0.0)) : 0.0;
    X[i * SLD + j] = 0.0;
    X[i * SLD + j + 1] = 0.0;
  }
  if (tid == 0) fail = 0;
  GPX_DIAG_STAMP(0);
  __syncthreads();
  GPX_DIAG_STAMP(1);
#pragma unroll 1
  for (int p = 0; p < np; p++) {
    const int c0 = PB * p;
    GPX_DIAG_STAMP(2 + 4 * p);
    if (warp == 0) {
      // lanes l and l + 16 carry the same row (identical values, benign duplicate stores)
      const int r = lane & 15;
      double a[PB];
#pragma unroll
      for (int j = 0; j < PB; j++) a[j] = S[(c0 + r) * SLD + c0 + j];
      double dg = S[(c0 + r) * SLD + c0 + r];    // running diagonal entry of this row
      int bad = 0;
      if (lane == 0) LB[LBS + PB] = dg;          // pivot of step 0 (step k reads the slot of buffer (k - 1) & 1)
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PB; k++) {
        double* cur = LB + (k & 1) * LBS;
        const double d = LB[((k + 1) & 1) * LBS + PB];
        if (!(d > 0.0) && bad == 0 && c0 + k < jb) bad = j0 + c0 + k + 1;      // also catches NaN
        const double rd = rsqrt_fast(d);
        const double l = ((r == k) ? dg : a[k]) * rd;
        a[k] = l;
        dg = fma(-l, l, dg);
        cur[r] = l;
        if (r == k + 1) cur[PB] = dg;
        if (lane == k) rdg[c0 + k] = rd;
        __syncwarp();
#pragma unroll
        for (int j = k + 1; j < PB; j++) a[j] = fma(-l, cur[j], a[j]);
      }
#pragma unroll
      for (int j = 0; j < PB; j++) S[(c0 + r) * SLD + c0 + j] = (j <= r) ? a[j] : 0.0;
      if (lane == 0 && bad != 0 && fail == 0) fail = bad;
      __syncwarp();
      GPX_DIAG_STAMP(3 + 4 * p);
    }
    __syncthreads();
    GPX_DIAG_STAMP(4 + 4 * p);
    const int n = nr - c0 - PB;        // live rows below the panel
    if (n > 0) {
      // panel: P L_D^T = A21 by forward substitution, one thread per row (right-looking, so the FMAs of a step are
      // independent).  Substitution, not a product with the sub-block inverse: for ill-conditioned K(z, z) the product
      // doubles the error of the unwhitened C3-shape parity test (1.5e-8 instead of 7e-9 on dL/dq_sqrt).
      if (tid < n) {
        double* srow = S + (c0 + PB + tid) * SLD + c0;
        double a[PB];
#pragma unroll
        for (int c = 0; c < PB; c++) a[c] = srow[c];
#pragma unroll
        for (int c = 0; c < PB; c++) {
          a[c] *= rdg[c0 + c];
#pragma unroll
          for (int c2 = c + 1; c2 < PB; c2++) a[c2] = fma(-a[c], S[(c0 + c2) * SLD + c0 + c], a[c2]);
        }
#pragma unroll
        for (int c = 0; c < PB; c++) srow[c] = a[c];
      }
      __syncthreads();
      GPX_DIAG_STAMP(5 + 4 * p);
      // trailing update (lower triangle): A22 -= P P^T, thread (ty, tx) owns the elements (ty + 16 a, tx + 16 c)
      double acc[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) acc[a][c] = 0.0;
      const double* prow = S + (c0 + PB + ty) * SLD + c0;
      const double* pcol = S + (c0 + PB + tx) * SLD + c0;
#pragma unroll
      for (int t0 = 0; t0 < PB; t0 += 4) {
        double pr[3][4], pj[3][4];
#pragma unroll
        for (int a = 0; a < 3; a++)
          if (16 * a < n) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              pr[a][u] = prow[16 * a * SLD + t0 + u];
              pj[a][u] = pcol[16 * a * SLD + t0 + u];
            }
          }
#pragma unroll
        for (int a = 0; a < 3; a++)
          if (16 * a < n) {
#pragma unroll
            for (int c = 0; c <= a; c++)
#pragma unroll
              for (int u = 0; u < 4; u++) acc[a][c] = fma(pr[a][u], pj[c][u], acc[a][c]);
          }
      }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c <= a; c++)
          if (16 * a < n && (c < a || tx <= ty)) S[(c0 + PB + ty + 16 * a) * SLD + c0 + PB + tx + 16 * c] -= acc[a][c];
      __syncthreads();
    }
  }
  GPX_DIAG_STAMP(18);
#pragma unroll 4
  for (int i = warp; i < jb; i += 8) {
    const int j = 2 * lane;
    const double v0 = (j <= i) ? S[i * SLD + j] : 0.0, v1 = (j + 1 <= i) ? S[i * SLD + j + 1] : 0.0;
    if (vecA && j + 1 < jb) *reinterpret_cast<double2*>(Ab + (long long)i * lda + j) = make_double2(v0, v1);
    else {
      if (j < jb) Ab[(long long)i * lda + j] = v0;
      if (j + 1 < jb) Ab[(long long)i * lda + j + 1] = v1;
    }
  }
  GPX_DIAG_STAMP(19);
  // ---- inverse, by substitution throughout.  Diagonal 16 x 16 blocks: warp i inverts block i (column per lane,
  // right-looking); then block row i = 1, 2, 3:  L_ii X_ij = -sum_{k=j}^{i-1} L_ik X_kj  -- the right-hand side is a set
  // of dot products over all threads, the solve runs one thread per column.
  if (warp < np) {
    const int c0 = PB * warp, r = lane & 15;
    double x[PB];
#pragma unroll
    for (int i = 0; i < PB; i++) x[i] = (i == r) ? 1.0 : 0.0;
#pragma unroll
    for (int t = 0; t < PB; t++) {
      x[t] *= rdg[c0 + t];
#pragma unroll
      for (int i = t + 1; i < PB; i++) x[i] = fma(-S[(c0 + i) * SLD + c0 + t], x[t], x[i]);
    }
#pragma unroll
    for (int i = 0; i < PB; i++) X[(c0 + i) * SLD + c0 + r] = x[i];
  }
  __syncthreads();
  GPX_DIAG_STAMP(20);
#pragma unroll
  for (int i = 1; i < 4; i++) {
    if (i < np) {
      const double* lrow = S + (16 * i + ty) * SLD;
      double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int kk = 0; kk < 16 * i; kk++) {
        const double lv = lrow[kk];
#pragma unroll
        for (int m = 0; m < i; m++)
          if (kk >= 16 * m) acc[m] = fma(lv, X[kk * SLD + 16 * m + tx], acc[m]);
      }
#pragma unroll
      for (int m = 0; m < i; m++) X[(16 * i + ty) * SLD + 16 * m + tx] = -acc[m];
    }
    __syncthreads();
    if (i < np && tid < 16 * i) {
      double* xc = X + 16 * i * SLD + tid;
      double x[PB];
#pragma unroll
      for (int r = 0; r < PB; r++) x[r] = xc[r * SLD];
#pragma unroll
      for (int t = 0; t < PB; t++) {
        x[t] *= rdg[16 * i + t];
#pragma unroll
        for (int r = t + 1; r < PB; r++) x[r] = fma(-S[(16 * i + r) * SLD + 16 * i + t], x[t], x[r]);
      }
#pragma unroll
      for (int r = 0; r < PB; r++) xc[r * SLD] = x[r];
    }
    __syncthreads();
    GPX_DIAG_STAMP(20 + i);
  }
  if (Linv) {
    double* Xb = Linv + (long long)b * sI + (long long)j0 * ldi + j0;
    const bool vecX = ((ldi & 1) == 0) && ((((uintptr_t)Xb) & 15) == 0);
#pragma unroll 4
    for (int i = warp; i < jb; i += 8) {
      const int j = 2 * lane;
      const double v0 = (j <= i) ? X[i * SLD + j] : 0.0, v1 = (j + 1 <= i) ? X[i * SLD + j + 1] : 0.0;
      if (vecX && j + 1 < jb) *reinterpret_cast<double2*>(Xb + (long long)i * ldi + j) = make_double2(v0, v1);
      else {
        if (j < jb) Xb[(long long)i * ldi + j] = v0;
        if (j + 1 < jb) Xb[(long long)i * ldi + j + 1] = v1;
      }
    }
  }
  GPX_DIAG_STAMP(24);
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

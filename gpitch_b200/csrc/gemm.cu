#include "gemm.cuh"

namespace gpx {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int BM, int BN, bool TA, bool TB>
struct Tile {
  static constexpr int WM = BM / 2, WN = BN / 4;   // 8 warps as 2 (m) x 4 (n)
  static constexpr int MT = WM / 8, NT = WN / 8;
  // +4 padding makes every fragment read bank-conflict free for 64-bit accesses (ld == 4 mod 16 doubles).
  static constexpr int A_ROWS = TA ? BK : BM, A_LD = (TA ? BM : BK) + 4;
  static constexpr int B_ROWS = TB ? BN : BK, B_LD = (TB ? BK : BN) + 4;
  static constexpr int A_ELEMS = A_ROWS * A_LD, B_ELEMS = B_ROWS * B_LD;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS + BK;  // + k-weights
};

// Copy a [ROWS x COLS] tile (COLS contiguous) from global (origin r0,c0; bounds R,C; leading dim ld) into smem
// with leading dim LD.  Out-of-range elements are zero-filled (cp.async src-size 0 / 8).
template <int ROWS, int COLS, int LD>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, int ld, int r0, int c0,
                                          int R, int C, bool vec_ok) {
  constexpr int CH = COLS / 2;  // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < ROWS * CH; idx += NTHREADS) {
    int r = idx / CH, c = (idx - r * CH) * 2;
    int gr = r0 + r, gc = c0 + c;
    double* dst = s + r * LD + c;
    bool rv = gr < R;
    int n = rv ? (C - gc) : 0;  // valid elements in this chunk (<=0 none, 1, >=2 both)
    const double* src = g + (long long)(rv ? gr : 0) * ld + (n > 0 ? gc : 0);
    if (vec_ok) {
      cp_async16(dst, src, n >= 2 ? 16 : (n == 1 ? 8 : 0));
    } else {
      cp_async8(dst, src, n >= 1 ? 8 : 0);
      cp_async8(dst + 1, n >= 2 ? src + 1 : src, n >= 2 ? 8 : 0);
    }
  }
}

template <int BM, int BN, bool TA, bool TB, int STAGES>
__global__ void __launch_bounds__(NTHREADS, (BM * BN > 80 * 128) ? 1 : 2) gemm_kernel(const GemmArgs p) {
  using T = Tile<BM, BN, TA, TB>;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = p.M, N = p.N, K = p.K;
  if ((p.flags & GEMM_C_LOWER) && n0 > m0 + BM - 1) {  // tile strictly above the diagonal: not computed
    if (p.flags & GEMM_ZERO_UPPER) {
      double* Cz = p.C + (long long)b * p.sC;
      for (int idx = threadIdx.x; idx < BM * BN; idx += NTHREADS) {
        const int r = m0 + idx / BN, c = n0 + idx % BN;
        if (r < M && c < N) Cz[(long long)r * p.ldc + c] = 0.0;
      }
    }
    return;
  }

  // k range for this tile from the triangular-structure flags
  int kb = 0, ke = K;
  if (p.flags & GEMM_A_LOWER) ke = min(ke, m0 + BM);
  if (p.flags & GEMM_A_UPPER) kb = max(kb, m0);
  if (p.flags & GEMM_B_LOWER) kb = max(kb, n0);
  if (p.flags & GEMM_B_UPPER) ke = min(ke, n0 + BN);
  kb = (kb / BK) * BK;
  const int nk = (ke > kb) ? (ke - kb + BK - 1) / BK : 0;

  const double* Ag = p.A + (long long)b * p.sA;
  const double* Bg = p.B + (long long)b * p.sB;
  const double* Wg = p.kweight ? p.kweight + (long long)b * p.sKw : nullptr;
  const bool vecA = ((p.lda & 1) == 0) && ((((uintptr_t)Ag) & 15) == 0);
  const bool vecB = ((p.ldb & 1) == 0) && ((((uintptr_t)Bg) & 15) == 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * T::WM, wn0 = (warp & 3) * T::WN;

  double acc[T::MT][T::NT][2];
#pragma unroll
  for (int i = 0; i < T::MT; i++)
#pragma unroll
    for (int j = 0; j < T::NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int stage, int kt) {
    double* sA = smem + stage * T::STAGE_ELEMS;
    double* sB = sA + T::A_ELEMS;
    double* sW = sB + T::B_ELEMS;
    const int k0 = kb + kt * BK;
    if (TA) load_tile<BK, BM, T::A_LD>(sA, Ag, p.lda, k0, m0, K, M, vecA);
    else    load_tile<BM, BK, T::A_LD>(sA, Ag, p.lda, m0, k0, M, K, vecA);
    if (TB) load_tile<BN, BK, T::B_LD>(sB, Bg, p.ldb, n0, k0, N, K, vecB);
    else    load_tile<BK, BN, T::B_LD>(sB, Bg, p.ldb, k0, n0, K, N, vecB);
    if (Wg && threadIdx.x < BK) {
      int k = k0 + threadIdx.x;
      cp_async8(sW + threadIdx.x, Wg + (k < K ? k : 0), k < K ? 8 : 0);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * T::STAGE_ELEMS;
    const double* sB = sA + T::A_ELEMS;
    const double* sW = sB + T::B_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[T::MT], bf[T::NT];
#pragma unroll
      for (int i = 0; i < T::MT; i++)
        af[i] = TA ? sA[(kk + t) * T::A_LD + wm0 + i * 8 + g] : sA[(wm0 + i * 8 + g) * T::A_LD + kk + t];
#pragma unroll
      for (int j = 0; j < T::NT; j++)
        bf[j] = TB ? sB[(wn0 + j * 8 + g) * T::B_LD + kk + t] : sB[(kk + t) * T::B_LD + wn0 + j * 8 + g];
      if (Wg) {
        double w = sW[kk + t];
#pragma unroll
        for (int j = 0; j < T::NT; j++) bf[j] *= w;
      }
#pragma unroll
      for (int i = 0; i < T::MT; i++)
#pragma unroll
        for (int j = 0; j < T::NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue
  double* Cg = p.C + (long long)b * p.sC;
  const double alpha = p.alpha * (p.alpha_vec ? p.alpha_vec[b] : 1.0);
  const double beta = p.beta, gamma = p.gamma;
  const double* Aux = p.Aux ? p.Aux + (long long)b * p.sAux : nullptr;
  const double* cs = p.colscale ? p.colscale + (long long)b * p.sColscale : nullptr;
  const double* rv = p.rowvec ? p.rowvec + (long long)b * p.sRowvec : nullptr;
  const double* cv = p.colvec ? p.colvec + (long long)b * p.sColvec : nullptr;
  const bool mirror = (p.flags & GEMM_C_MIRROR) && (p.flags & GEMM_C_LOWER);
  const bool zero_upper = (p.flags & GEMM_ZERO_UPPER) && (p.flags & GEMM_C_LOWER);
#pragma unroll
  for (int i = 0; i < T::MT; i++) {
    const int row = m0 + wm0 + i * 8 + g;
    if (row >= M) continue;
    const double rvv = rv ? rv[row] : 0.0;
#pragma unroll
    for (int j = 0; j < T::NT; j++) {
      const int col = n0 + wn0 + j * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = col + e;
        if (c >= N) continue;
        double v = alpha * acc[i][j][e];
        if (Aux) v += gamma * Aux[(long long)row * p.ldaux + c];
        if (cs) v *= cs[c];
        if (rv) v += rvv * cv[c];
        if (beta != 0.0) v += beta * Cg[(long long)row * p.ldc + c];
        if (zero_upper && c > row) v = 0.0;
        Cg[(long long)row * p.ldc + c] = v;
        if (mirror && c < row) Cg[(long long)c * p.ldc + row] = v;
      }
    }
  }
}

template <int BM, int BN, bool TA, bool TB>
static int launch_cfg(const GemmArgs& a, cudaStream_t st) {
  constexpr int STAGES = 3;
  using T = Tile<BM, BN, TA, TB>;
  size_t smem = (size_t)STAGES * T::STAGE_ELEMS * sizeof(double);
  auto kern = gemm_kernel<BM, BN, TA, TB, STAGES>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.batch);
  kern<<<grid, NTHREADS, smem, st>>>(a);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

template <int BM, int BN>
static int launch_trans(const GemmArgs& a, cudaStream_t st) {
  const bool ta = a.flags & GEMM_TRANS_A, tb = a.flags & GEMM_TRANS_B;
  if (!ta && !tb) return launch_cfg<BM, BN, false, false>(a, st);
  if (ta && !tb) return launch_cfg<BM, BN, true, false>(a, st);
  if (!ta && tb) return launch_cfg<BM, BN, false, true>(a, st);
  return launch_cfg<BM, BN, true, true>(a, st);
}

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return GPX_OK;
  if (a.K < 0 || a.batch > 65535) return GPX_ERR_ARG;
  // Row-tile height: 80 divides the M = 200 / 400 inducing sets of the named configs exactly (no padded rows);
  // 128 when it wastes fewer padded rows (M = 2048, 128, 256 ...).
  const int waste80 = (a.M + 79) / 80 * 80 - a.M, waste128 = (a.M + 127) / 128 * 128 - a.M;
  const bool narrow = a.N <= 64;
  if (waste128 < waste80) return narrow ? launch_trans<128, 64>(a, st) : launch_trans<128, 128>(a, st);
  return narrow ? launch_trans<80, 64>(a, st) : launch_trans<80, 128>(a, st);
}

}  // namespace gpx

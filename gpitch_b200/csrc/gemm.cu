#include "gemm.cuh"
#include <cstdlib>

namespace gpx {

constexpr int BK = 16;
constexpr int STAGES = 3;

template <int BM, int BN, int WGN, bool TA, bool TB>
struct Tile {
  static constexpr int WGM = 2;                     // warp grid: 2 (m) x WGN (n); WGN = 4 -> 256 threads, 2 -> 128
  static constexpr int NTH = 32 * WGM * WGN;
  static constexpr int WM = BM / WGM, WN = BN / WGN;
  static constexpr int MT = WM / 8, NT = WN / 8;
  // +4 padding makes every fragment read bank-conflict free for 64-bit accesses (ld == 4 mod 16 doubles).
  static constexpr int A_ROWS = TA ? BK : BM, A_COLS = TA ? BM : BK, A_LD = A_COLS + 4;
  static constexpr int B_ROWS = TB ? BN : BK, B_COLS = TB ? BK : BN, B_LD = B_COLS + 4;
  static constexpr int A_ELEMS = A_ROWS * A_LD, B_ELEMS = B_ROWS * B_LD;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS + BK;  // + k-weights
  static constexpr int A_NE = (A_ROWS * (A_COLS / 2) + NTH - 1) / NTH;  // 16-byte chunks per thread
  static constexpr int B_NE = (B_ROWS * (B_COLS / 2) + NTH - 1) / NTH;
};

// Generic (slow-path) tile copy: origin (r0, c0), bounds (R, C); out-of-range elements are zero-filled.  Used for
// the last, partial k-tile and for operands that are not 16-byte aligned.
template <int ROWS, int COLS, int LD, int NTH>
__device__ __forceinline__ void load_tile_generic(double* __restrict__ s, const double* __restrict__ g, int ld, int r0,
                                                  int c0, int R, int C, bool vec_ok) {
  constexpr int CH = COLS / 2;
  for (int idx = threadIdx.x; idx < ROWS * CH; idx += NTH) {
    int r = idx / CH, c = (idx - r * CH) * 2;
    int gr = r0 + r, gc = c0 + c;
    double* dst = s + r * LD + c;
    bool rv = gr < R;
    int n = rv ? (C - gc) : 0;
    const double* src = g + (long long)(rv ? gr : 0) * ld + (n > 0 ? gc : 0);
    if (vec_ok) {
      cp_async16(dst, src, n >= 2 ? 16 : (n == 1 ? 8 : 0));
    } else {
      cp_async8(dst, src, n >= 1 ? 8 : 0);
      cp_async8(dst + 1, n >= 2 ? src + 1 : src, n >= 2 ? 8 : 0);
    }
  }
}

// Per-thread copy plan of one operand tile: chunk e of this thread lives at smem offset so[e] and global pointer
// gp[e] (for the first k-tile); every further k-tile just bumps the pointers.  nb[e] = bytes valid w.r.t. the
// non-k dimension (0 / 8 / 16).  Removes all index arithmetic and bounds checks from the main loop.
template <int NE>
struct Plan {
  const double* gp[NE];
  int so[NE];
  int nb[NE];
};

template <int ROWS, int COLS, int LD, int NE, bool K_IS_ROW, int NTH>
__device__ __forceinline__ void make_plan(Plan<NE>& p, const double* g, int ld, int r0, int c0, int R, int C) {
  constexpr int CH = COLS / 2;
#pragma unroll
  for (int e = 0; e < NE; e++) {
    const int idx = threadIdx.x + e * NTH;
    const int r = idx / CH, c = (idx - r * CH) * 2;
    const bool in_tile = idx < ROWS * CH;
    p.so[e] = in_tile ? r * LD + c : -1;
    int n;
    long long off;
    if (K_IS_ROW) {   // rows walk k (always valid on the fast path); columns are the m / n dimension
      n = C - (c0 + c);
      off = (long long)(r0 + r) * ld + (n > 0 ? c0 + c : 0);
    } else {          // rows are the m / n dimension; columns walk k
      n = (r0 + r < R) ? 2 : 0;
      off = (long long)(n > 0 ? r0 + r : 0) * ld + c0 + c;
    }
    p.nb[e] = in_tile ? (n >= 2 ? 16 : (n == 1 ? 8 : 0)) : 0;
    p.gp[e] = g + off;
  }
}

template <int NE>
__device__ __forceinline__ void issue_plan(Plan<NE>& p, double* s, long long step) {
#pragma unroll
  for (int e = 0; e < NE; e++) {
    if (p.so[e] >= 0) cp_async16(s + p.so[e], p.nb[e] ? p.gp[e] : p.gp[0], p.nb[e]);
    p.gp[e] += step;
  }
}

template <int BM, int BN, int WGN, bool TA, bool TB, bool HAS_W>
__global__ void __launch_bounds__(32 * 2 * WGN, (WGN == 2) ? 3 : ((BM * BN > 80 * 128) ? 1 : 2)) gemm_kernel(const GemmArgs p) {
  using T = Tile<BM, BN, WGN, TA, TB>;
  constexpr int NTH = T::NTH;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = p.M, N = p.N, K = p.K;
  const bool c_lower = p.flags & GEMM_C_LOWER;
  if (c_lower && n0 > m0 + BM - 1) {  // tile strictly above the diagonal: not computed
    if (p.flags & GEMM_ZERO_UPPER) {
      double* Cz = p.C + (long long)b * p.sC;
      for (int idx = threadIdx.x; idx < BM * BN; idx += NTH) {
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
  const double* Wg = HAS_W ? p.kweight + (long long)b * p.sKw : nullptr;
  const bool vecA = ((p.lda & 1) == 0) && ((((uintptr_t)Ag) & 15) == 0);
  const bool vecB = ((p.ldb & 1) == 0) && ((((uintptr_t)Bg) & 15) == 0);
  // k-origins must keep 16-byte alignment on the fast path: kb is a multiple of 16 and m0/n0 of 8 -> always even.
  const bool fast = vecA && vecB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / WGN) * T::WM, wn0 = (warp % WGN) * T::WN;
#define AROW(i) (wm0 + (i) * 8)
#define BCOL(j) (wn0 + (j) * 8)
  // NB a predicated-off DMMA still occupies the FP64 pipe (measured: masking individual m8n8 sub-tiles buys
  // nothing, tools/gemm_bench.py), so structural zeros are skipped only by warp-uniform BRANCHES: a warp whose
  // whole tile lies above the diagonal of a lower-only output does no math at all ...
  const bool warp_dead = c_lower && (n0 + wn0 > m0 + wm0 + T::WM - 1);

  double acc[T::MT][T::NT][2];
#pragma unroll
  for (int i = 0; i < T::MT; i++)
#pragma unroll
    for (int j = 0; j < T::NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  Plan<T::A_NE> pa;
  Plan<T::B_NE> pb;
  if (fast) {
    if (TA) make_plan<BK, BM, T::A_LD, T::A_NE, true, NTH>(pa, Ag, p.lda, kb, m0, K, M);
    else    make_plan<BM, BK, T::A_LD, T::A_NE, false, NTH>(pa, Ag, p.lda, m0, kb, M, K);
    if (TB) make_plan<BN, BK, T::B_LD, T::B_NE, false, NTH>(pb, Bg, p.ldb, n0, kb, N, K);
    else    make_plan<BK, BN, T::B_LD, T::B_NE, true, NTH>(pb, Bg, p.ldb, kb, n0, K, N);
  }
  const long long stepA = TA ? (long long)BK * p.lda : BK;
  const long long stepB = TB ? BK : (long long)BK * p.ldb;
  int k_next = kb;  // k origin of the next tile to be loaded

  auto load_stage = [&](int stage) {
    double* sA = smem + stage * T::STAGE_ELEMS;
    double* sB = sA + T::A_ELEMS;
    double* sW = sB + T::B_ELEMS;
    const int k0 = k_next;
    if (fast && k0 + BK <= K) {
      issue_plan<T::A_NE>(pa, sA, stepA);
      issue_plan<T::B_NE>(pb, sB, stepB);
    } else {
      if (TA) load_tile_generic<BK, BM, T::A_LD, NTH>(sA, Ag, p.lda, k0, m0, K, M, vecA);
      else    load_tile_generic<BM, BK, T::A_LD, NTH>(sA, Ag, p.lda, m0, k0, M, K, vecA);
      if (TB) load_tile_generic<BN, BK, T::B_LD, NTH>(sB, Bg, p.ldb, n0, k0, N, K, vecB);
      else    load_tile_generic<BK, BN, T::B_LD, NTH>(sB, Bg, p.ldb, k0, n0, K, N, vecB);
    }
    if (HAS_W && threadIdx.x < BK) {
      int k = k0 + threadIdx.x;
      cp_async8(sW + threadIdx.x, Wg + (k < K ? k : 0), k < K ? 8 : 0);
    }
    k_next += BK;
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s);
    cp_async_commit();
  }

  const int tri_flags = p.flags & (GEMM_A_LOWER | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_B_UPPER);

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < nk) load_stage((kt + STAGES - 1) % STAGES);
    cp_async_commit();
    const double* sA = smem + (kt % STAGES) * T::STAGE_ELEMS;
    const double* sB = sA + T::A_ELEMS;
    const double* sW = sB + T::B_ELEMS;
    const int k0 = kb + kt * BK;
    // ... and a warp skips a whole k-tile when a triangular operand makes its A rows / B columns zero there.
    if (warp_dead) continue;
    if (tri_flags) {
      const int row_lo = m0 + wm0, col_lo = n0 + wn0;
      if ((tri_flags & GEMM_A_LOWER) && k0 > row_lo + T::WM - 1) continue;      // k > every row of the warp tile
      if ((tri_flags & GEMM_A_UPPER) && k0 + BK - 1 < row_lo) continue;
      if ((tri_flags & GEMM_B_LOWER) && k0 + BK - 1 < col_lo) continue;
      if ((tri_flags & GEMM_B_UPPER) && k0 > col_lo + T::WN - 1) continue;
    }
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[T::MT], bf[T::NT];
#pragma unroll
      for (int i = 0; i < T::MT; i++)
        af[i] = TA ? sA[(kk + t) * T::A_LD + AROW(i) + g] : sA[(AROW(i) + g) * T::A_LD + kk + t];
#pragma unroll
      for (int j = 0; j < T::NT; j++)
        bf[j] = TB ? sB[(BCOL(j) + g) * T::B_LD + kk + t] : sB[(kk + t) * T::B_LD + BCOL(j) + g];
      if (HAS_W) {
        const double w = sW[kk + t];
#pragma unroll
        for (int j = 0; j < T::NT; j++) bf[j] *= w;
      }
#pragma unroll
      for (int i = 0; i < T::MT; i++)
#pragma unroll
        for (int j = 0; j < T::NT; j++)
          dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  if (warp_dead) {
    if (p.flags & GEMM_ZERO_UPPER) {   // this warp's tile lies entirely above the diagonal
      double* Cz = p.C + (long long)b * p.sC;
      for (int idx = lane; idx < T::WM * T::WN; idx += 32) {
        const int r = m0 + wm0 + idx / T::WN, c = n0 + wn0 + idx % T::WN;
        if (r < M && c < N) Cz[(long long)r * p.ldc + c] = 0.0;
      }
    }
    return;
  }

  // ---- epilogue
  double* Cg = p.C + (long long)b * p.sC;
  const double alpha = p.alpha * (p.alpha_vec ? p.alpha_vec[b] : 1.0);
  const double beta = p.beta, gamma = p.gamma * (p.gamma_vec ? p.gamma_vec[b] : 1.0);
  const double* Aux = p.Aux ? p.Aux + (long long)b * p.sAux : nullptr;
  const double* cs = p.colscale ? p.colscale + (long long)b * p.sColscale : nullptr;
  const double* rv = p.rowvec ? p.rowvec + (long long)b * p.sRowvec : nullptr;
  const double* cv = p.colvec ? p.colvec + (long long)b * p.sColvec : nullptr;
  const bool mirror = (p.flags & GEMM_C_MIRROR) && c_lower;
  const bool zero_upper = (p.flags & GEMM_ZERO_UPPER) && c_lower;
  const bool vecC = ((p.ldc & 1) == 0) && ((((uintptr_t)Cg) & 15) == 0) && !c_lower;
#pragma unroll
  for (int i = 0; i < T::MT; i++) {
    const int row = m0 + AROW(i) + g;
    if (row >= M) continue;
    const double rvv = rv ? rv[row] : 0.0;
#pragma unroll
    for (int j = 0; j < T::NT; j++) {
      const int col = n0 + BCOL(j) + 2 * t;
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = col + e;
        double x = alpha * acc[i][j][e];
        if (c < N) {
          if (Aux) x += gamma * Aux[(long long)row * p.ldaux + c];
          if (cs) x *= cs[c];
          if (rv) x += rvv * cv[c];
          if (beta != 0.0) x += beta * Cg[(long long)row * p.ldc + c];
        }
        v[e] = x;
      }
      if (vecC && col + 1 < N) {
        *reinterpret_cast<double2*>(Cg + (long long)row * p.ldc + col) = make_double2(v[0], v[1]);
      } else {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = col + e;
          if (c >= N) continue;
          if (c_lower && c > row) {        // above the diagonal of a lower-only result
            if (zero_upper) Cg[(long long)row * p.ldc + c] = 0.0;
            continue;
          }
          Cg[(long long)row * p.ldc + c] = v[e];
          if (mirror && c < row) Cg[(long long)c * p.ldc + row] = v[e];
        }
      }
    }
  }
}

template <int BM, int BN, int WGN, bool TA, bool TB, bool HAS_W>
static int launch_cfg(const GemmArgs& a, cudaStream_t st) {
  using T = Tile<BM, BN, WGN, TA, TB>;
  size_t smem = (size_t)STAGES * T::STAGE_ELEMS * sizeof(double);
  auto kern = gemm_kernel<BM, BN, WGN, TA, TB, HAS_W>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: set every time
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.batch);
  kern<<<grid, T::NTH, smem, st>>>(a);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

template <int BM, int BN, int WGN>
static int launch_trans(const GemmArgs& a, cudaStream_t st) {
  const bool ta = a.flags & GEMM_TRANS_A, tb = a.flags & GEMM_TRANS_B;
  if (a.kweight) {  // k-weights: instantiated for the A diag(w) B^T form only (operands stored [M,K] and [N,K])
    if (ta || !tb) return GPX_ERR_ARG;
    return launch_cfg<BM, BN, WGN, false, true, true>(a, st);
  }
  if (!ta && !tb) return launch_cfg<BM, BN, WGN, false, false, false>(a, st);
  if (ta && !tb) return launch_cfg<BM, BN, WGN, true, false, false>(a, st);
  if (!ta && tb) return launch_cfg<BM, BN, WGN, false, true, false>(a, st);
  return launch_cfg<BM, BN, WGN, true, true, false>(a, st);
}

static int launch_gemm_cpasync(const GemmArgs& a, cudaStream_t st);

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return GPX_OK;
  if (a.K < 0) return GPX_ERR_ARG;
  const int rc = launch_gemm_tma(a, st);      // Blackwell data path (TMA + mbarrier ring) whenever alignment allows
  if (rc != 1) return rc;
  // cp.async kernel: grid.z carries the batch -> slices of at most 65535 entries
  for (int b0 = 0; b0 < a.batch; b0 += 65535) {
    GemmArgs s = a;
    s.batch = (a.batch - b0 < 65535) ? a.batch - b0 : 65535;
    s.A += (long long)b0 * a.sA; s.B += (long long)b0 * a.sB; s.C += (long long)b0 * a.sC;
    if (a.alpha_vec) s.alpha_vec += b0;
    if (a.kweight) s.kweight += (long long)b0 * a.sKw;
    if (a.Aux) s.Aux += (long long)b0 * a.sAux;
    if (a.colscale) s.colscale += (long long)b0 * a.sColscale;
    if (a.rowvec) s.rowvec += (long long)b0 * a.sRowvec;
    if (a.colvec) s.colvec += (long long)b0 * a.sColvec;
    const int r = launch_gemm_cpasync(s, st);
    if (r != GPX_OK) return r;
  }
  return GPX_OK;
}

static int launch_gemm_cpasync(const GemmArgs& a, cudaStream_t st) {
  // Row-tile height: 80 divides the M = 200 / 400 inducing sets of the named configs exactly (no padded rows);
  // 128 when it wastes fewer padded rows (M = 2048, 128, 256 ...).
  const int waste80 = (a.M + 79) / 80 * 80 - a.M, waste128 = (a.M + 127) / 128 * 128 - a.M;
  static const int narrow_max = getenv("GPX_NARROW_MAXN") ? atoi(getenv("GPX_NARROW_MAXN")) : 512;
  static const int sq80 = getenv("GPX_SQ80") ? atoi(getenv("GPX_SQ80")) : 1;
  const bool tri = a.flags & (GEMM_A_LOWER | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_B_UPPER);
  const bool syrk = (a.flags & GEMM_C_LOWER) && !tri;
  if (waste128 < waste80) return (a.N <= 64) ? launch_trans<128, 64, 4>(a, st) : launch_trans<128, 128, 4>(a, st);
  if (a.N <= 64) return launch_trans<80, 64, 4>(a, st);
  // M x M x M products (N <= 512) and lower-only SYRK outputs: square 80 x 80 tiles on 4 warps -- 400 = 5 x 80 leaves no
  // padded columns and the tiles hug the diagonal; 64-wide 8-warp tiles otherwise
  if (a.N <= narrow_max || syrk) {
    if (sq80 && ((a.N + 79) / 80 * 80 - a.N) <= ((a.N + 63) / 64 * 64 - a.N)) return launch_trans<80, 80, 2>(a, st);
    return launch_trans<80, 64, 4>(a, st);
  }
  // Large N: 4-warp CTAs (3 per SM, three independent barrier domains) beat the 8-warp 80 x 128 tile by 5-10 % on
  // B200 (tools/gemm_bench.py: dense 8.18 vs 9.03 ms); 80 x 64 for op(A) = A, 80 x 80 for op(A) = A^T.
  static const int bigcfg = getenv("GPX_BIGCFG") ? atoi(getenv("GPX_BIGCFG")) : -1;   // tile experiments
  if (bigcfg == 0) return launch_trans<80, 128, 4>(a, st);
  if (bigcfg == 1) return launch_trans<80, 64, 2>(a, st);
  if (bigcfg == 2) return launch_trans<80, 80, 2>(a, st);
  return (a.flags & GEMM_TRANS_A) ? launch_trans<80, 80, 2>(a, st) : launch_trans<80, 64, 2>(a, st);
}

#undef AROW
#undef BCOL
}  // namespace gpx

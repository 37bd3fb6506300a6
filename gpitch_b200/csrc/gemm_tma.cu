// Batched fp64 GEMM, Blackwell data path: operand tiles arrive by TMA (cp.async.bulk.tensor -> SASS UTMALDG) into a
// ring of shared-memory stages guarded by mbarriers (SASS SYNCS); the math stays on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> DMMA.8x8x4; tcgen05 has no f64 kind).  Same contract as gemm.cu's cp.async kernel
// (launch_gemm picks this path whenever the operands meet TMA's 16-byte alignment rules), replacing
// tf.matmul / tf.matrix_triangular_solve of gpitch/sgpr_ss.py:48-53 and GPflow conditional() (gpitch/pdgp.py:147-155).
//
// Pipeline.  No CTA-wide barrier in the main loop: every consumer warp waits on the stage's `full` mbarrier, feeds
// DMMA from shared memory and arrives on the stage's `empty` mbarrier; one elected lane (warp 0) re-arms the stage
// that was released one k-tile earlier (slack of one tile, prefetch distance STAGES-1) with expect_tx + TMA.
// No thread issues a copy instruction or computes a global address per element; out-of-range rows / columns / k
// are zero-filled by the TMA unit.
//
// Shared-memory layouts (conflict-free 64-bit fragment reads without padding):
//   k-contiguous operand (A stored [M,K]; B stored [N,K]): one box {16 k, rows} per stage with the 128-byte swizzle;
//     MMA lane g reads row 2 (g & 3) + (g >> 2) of each 8-row block (a row permutation inside the block, undone in
//     the epilogue), so the 16 lanes of a half warp touch 16 distinct 8-byte slots.
//   m/n-contiguous operand (A stored [K,M]; B stored [K,N]): boxes of {8 columns, 16 k} (64-byte rows, 64-byte
//     swizzle), one per 8-column MMA block; MMA lane g reads column (g & 1) + 4 ((g >> 1) & 1) + 2 (g >> 2) of the
//     block -- again a permutation inside the block under which every half warp covers all 16 slots, and which keeps
//     the two accumulator columns of a lane adjacent (16-byte stores).
//   (tests/test_gemm_tma_layout.py replays both layouts, the swizzles and the lane maps on the CPU.)
//
// Triangular structure at MMA granularity: besides skipping whole k-tiles per warp, k-tiles that straddle the
// diagonal of a triangular A run a row-block loop with exact k-ranges (real loops, not predicated DMMA, which would
// still occupy the pipe), and diagonal CTA tiles of a lower-only (SYRK) output run statically pruned copies of the
// main loop; a warp's MMA blocks are interleaved over the CTA tile so that this pruning removes the same amount of work
// from every warp.  Executed / algorithmic work: TRMM 1.02 (cp.async kernel: 1.12-1.2), SYRK 1.02 (1.2) at M = 400.
#include "gemm.cuh"
#include <cuda.h>
#include <cstdlib>

namespace gpx {
std::atomic<unsigned long long> g_gemm_tma_launches{0};   // launches that took the TMA path (tests assert it is the one that runs)
namespace {

constexpr int TBK = 16;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int BM, int BN, int WGN, bool TA, bool TB, bool HAS_W, int STAGES>
struct TCfg {
  static constexpr int WGM = 2;
  static constexpr int NTH = 32 * WGM * WGN;
  static constexpr int WM = BM / WGM, WN = BN / WGN;
  static constexpr int MT = WM / 8, NT = WN / 8;
  static constexpr int A_ELEMS = BM * TBK, B_ELEMS = BN * TBK;          // 8-row / 8-column blocks of 128 doubles
  static constexpr int AS = WGM * 128, BS = WGN * 128;                  // a warp's blocks are interleaved (see kernel)
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;                 // multiple of 128 doubles = 1024 bytes
  static constexpr unsigned TX_BYTES = (unsigned)(STAGE_ELEMS + (HAS_W ? TBK : 0)) * 8u;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * 8 + (size_t)STAGES * TBK * 8 + 2 * STAGES * 8 + 1024;
  // split-K: a CTA's accumulators are parked in its (dead) operand ring for the cluster leader to collect
  static constexpr bool CAN_SPLIT = MT * NT * 2 * NTH <= STAGES * STAGE_ELEMS;
};

// Offset (doubles) of k4-step kk inside an operand tile, relative to the lane's base: k-contiguous tiles are rows of
// 16 k with the 128-byte swizzle (chunk ^= row & 7 -> kk ^ y2), m/n-contiguous tiles are 8-column blocks of 16 rows of
// 64 bytes with the 64-byte swizzle (chunk ^= (k >> 1) & 3 -> z2 ^ (kk & 4)).
template <bool KCONTIG>
__device__ __forceinline__ int koff(int kk, int y2, int z2) {
  return KCONTIG ? (kk ^ y2) : kk * 8 + (z2 ^ (kk & 4));
}

// One full k-tile of the warp's blocks.  PRUNE (lower-only output, CTA tile on the diagonal): 1 = blocks with j > i,
// 2 = blocks with j >= i are never needed (compile-time pruning: no predicate, no pipe slot).
template <class T, bool TA, bool TB, bool HAS_W, int PRUNE>
__device__ __forceinline__ void tile_full(double (&acc)[T::MT][T::NT][2], const double* __restrict__ sA,
                                          const double* __restrict__ sB, const double* __restrict__ sW, int a_base,
                                          int b_base, int y2, int z2, int t) {
#pragma unroll
  for (int kk = 0; kk < TBK; kk += 4) {
    const int ak = koff<!TA>(kk, y2, z2), bk = koff<TB>(kk, y2, z2);
    double af[T::MT], bf[T::NT];
#pragma unroll
    for (int i = 0; i < T::MT; i++) af[i] = sA[a_base + i * T::AS + ak];
#pragma unroll
    for (int j = 0; j < T::NT; j++) bf[j] = sB[b_base + j * T::BS + bk];
    if (HAS_W) {
      const double w = sW[kk + t];
#pragma unroll
      for (int j = 0; j < T::NT; j++) bf[j] *= w;
    }
#pragma unroll
    for (int i = 0; i < T::MT; i++)
#pragma unroll
      for (int j = 0; j < T::NT; j++) {
        if ((PRUNE == 1 && j > i) || (PRUNE == 2 && j >= i)) continue;
        dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
  }
}

template <int BM, int BN, int WGN, bool TA, bool TB, bool HAS_W, int STAGES>
__global__ void __launch_bounds__(32 * 2 * WGN, (WGN == 2) ? 3 : ((BM * BN > 80 * 128) ? 1 : 2))
    gemm_tma_kernel(const GemmArgs p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const int n_nt, const int n_mt, const int one_box, const int splitk) {
  using T = TCfg<BM, BN, WGN, TA, TB, HAS_W, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms (pointer arithmetic on the __shared__ array keeps LDS addressing)
  double* smem = reinterpret_cast<double*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  double* sWall = smem + STAGES * T::STAGE_ELEMS;
  uint64_t* full = reinterpret_cast<uint64_t*>(sWall + STAGES * TBK);
  uint64_t* empty = full + STAGES;

  // tile coordinates: n fastest, then m (heaviest row tiles first when A is lower triangular), then batch
  // split-K (few tiles, long K: the A A^T / weighted-SYRK launches of a single window): the `splitk` CTAs of a thread-block
  // cluster share one output tile, each takes a contiguous slice of the k-tiles, the leader collects the partial
  // accumulators through distributed shared memory and runs the epilogue (no workspace, fixed summation order)
  int idx = blockIdx.x / splitk;
  const int krank = blockIdx.x - idx * splitk;
  const int nt_i = idx % n_nt;
  idx /= n_nt;
  int mt_i = idx % n_mt;
  const int b = idx / n_mt;
  if (p.flags & GEMM_A_LOWER) mt_i = n_mt - 1 - mt_i;
  const int m0 = mt_i * BM, n0 = nt_i * BN;
  const int M = p.M, N = p.N, K = p.K;
  const bool c_lower = p.flags & GEMM_C_LOWER;
  if (c_lower && n0 > m0 + BM - 1) {  // tile strictly above the diagonal: not computed
    if (p.flags & GEMM_ZERO_UPPER) {
      double* Cz = p.C + (long long)b * p.sC;
      for (int e = threadIdx.x; e < BM * BN; e += T::NTH) {
        const int r = m0 + e / BN, c = n0 + e % BN;
        if (r < M && c < N) Cz[(long long)r * p.ldc + c] = 0.0;
      }
    }
    return;
  }

  int kb = 0, ke = K;
  if (p.flags & GEMM_A_LOWER) ke = min(ke, m0 + BM);
  if (p.flags & GEMM_A_UPPER) kb = max(kb, m0);
  if (p.flags & GEMM_B_LOWER) kb = max(kb, n0);
  if (p.flags & GEMM_B_UPPER) ke = min(ke, n0 + BN);
  kb = (kb / TBK) * TBK;
  int nk = (ke > kb) ? (ke - kb + TBK - 1) / TBK : 0;
  if (splitk > 1) {
    const int per = (nk + splitk - 1) / splitk;
    const int k_lo = min(nk, krank * per), k_hi = min(nk, k_lo + per);
    kb += k_lo * TBK;
    nk = k_hi - k_lo;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int rho = 2 * (g & 3) + (g >> 2);            // row of an 8-row block read by lane group g (k-contiguous operands)
  const int y2 = 2 * ((t >> 1) ^ rho);               // 128-byte swizzle term: 16-byte chunk index ^= row & 7
  const int nu = (g & 1) + 4 * ((g >> 1) & 1) + 2 * (g >> 2);   // column of an 8-column block (m/n-contiguous operands)
  const int z2 = 2 * ((nu >> 1) ^ (t >> 1));         // 64-byte swizzle term: chunk index ^= (k >> 1) & 3
  // Warp (wr, wc) of the WGM x WGN warp grid owns the 8-row blocks wr, wr + WGM, ... and the 8-column blocks wc, wc + WGN,
  // ... of the CTA tile (interleaved, not contiguous): on tiles that touch the diagonal of a triangular A or of a
  // lower-only C every warp then carries (almost) the same number of live MMA blocks, so no warp idles at the stage
  // barriers while another one finishes (contiguous sub-tiles: 25 / 15 / 15 / 0 blocks on a SYRK diagonal tile,
  // interleaved: 15 / 15 / 15 / 10).
  const int wr = warp / WGN, wc = warp % WGN;
  const int a_base = TA ? wr * 128 + t * 8 + (nu & 1) : (8 * wr + rho) * TBK + (t & 1);
  const int b_base = TB ? (8 * wc + rho) * TBK + (t & 1) : wc * 128 + t * 8 + (nu & 1);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * WGN);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int bA = p.sA ? b : 0, bB = p.sB ? b : 0;
  const double* Wg = HAS_W ? p.kweight + (long long)b * p.sKw : nullptr;
  auto issue = [&](int L) {     // k-tile L of this CTA -> stage L % STAGES (one thread)
    const int st = L % STAGES, k0 = kb + L * TBK;
    uint64_t* fb = &full[st];
    double* sA = smem + st * T::STAGE_ELEMS;
    double* sB = sA + T::A_ELEMS;
    mbar_expect_tx(fb, T::TX_BYTES);
    // m/n-contiguous tiles: ONE 4-d box {8 columns, 16 k, B/8 column blocks} when the column count is a multiple of 8
    // (the 4-d view splits the contiguous dimension into (c % 8, c / 8)), else one 3-d box per 8-column block
    if (!TA) tma_load_3d(sA, &mapA, fb, k0, m0, bA);
    else if (one_box & 1) tma_load_4d(sA, &mapA, fb, 0, k0, m0 >> 3, bA);
    else {
#pragma unroll 1
      for (int j = 0; j < BM / 8; j++) tma_load_3d(sA + j * 128, &mapA, fb, m0 + 8 * j, k0, bA);
    }
    if (TB) tma_load_3d(sB, &mapB, fb, k0, n0, bB);
    else if (one_box & 2) tma_load_4d(sB, &mapB, fb, 0, k0, n0 >> 3, bB);
    else {
#pragma unroll 1
      for (int j = 0; j < BN / 8; j++) tma_load_3d(sB + j * 128, &mapB, fb, n0 + 8 * j, k0, bB);
    }
    if (HAS_W) bulk_load_1d(sWall + st * TBK, Wg + k0, TBK * 8, fb);
  };
  if (threadIdx.x == 0) {
    for (int L = 0; L < STAGES - 1 && L < nk; L++) issue(L);
  }

  // lower-only output, square CTA tile on the diagonal: block (i, j) of this warp = (i WGM + wr, j WGN + wc) is needed
  // iff j WGN + wc <= i WGM + wr  ->  j <= i (wc <= wr) or j < i (wc > wr) when the warp grid is square
  const int prune = (c_lower && BM == BN && T::WGM == WGN && m0 == n0) ? (wc <= wr ? 1 : 2) : 0;
  const int tri_flags = p.flags & (GEMM_A_LOWER | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_B_UPPER);
  const int row_lo = m0 + 8 * wr, row_hi = m0 + 8 * ((T::MT - 1) * T::WGM + wr) + 7;      // first / last row of this warp
  const int col_lo = n0 + 8 * wc, col_hi = n0 + 8 * ((T::NT - 1) * WGN + wc) + 7;

  double acc[T::MT][T::NT][2];
#pragma unroll
  for (int i = 0; i < T::MT; i++)
#pragma unroll
    for (int j = 0; j < T::NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < nk; kt++) {
    const int st = kt % STAGES;
    mbar_wait(&full[st], (unsigned)((kt / STAGES) & 1));
    const double* sA = smem + st * T::STAGE_ELEMS;
    const double* sB = sA + T::A_ELEMS;
    const double* sW = sWall + st * TBK;
    const int k0 = kb + kt * TBK;
    // 0 = skip, 1 = every k4-step of every block is live, 2 / 3 = the k-tile straddles the diagonal of a triangular A / B
    int mode = 1;
    if (tri_flags) {
      const int kend = k0 + TBK - 1;
      int ma = 1, mb = 1;
      if (tri_flags & GEMM_A_LOWER) {            // op(A)[m,k] = 0 for k > m: step at k live for block rows r0.. iff k <= r0 + 7
        if (k0 > row_hi) ma = 0;
        else if (k0 + 5 > row_lo) ma = 2;        // (k0 + 12 <= row_lo + 7: the last step of the first block is live)
      }
      if (tri_flags & GEMM_A_UPPER) {            // op(A)[m,k] = 0 for k < m: step at k live iff k + 3 >= r0
        if (kend < row_lo) ma = 0;
        else if (k0 + 3 < row_hi - 7) ma = 2;
      }
      if (tri_flags & GEMM_B_LOWER) {            // op(B)[k,n] = 0 for k < n: step at k live for block columns c0.. iff k + 3 >= c0
        if (kend < col_lo) mb = 0;
        else if (k0 + 3 < col_hi - 7) mb = 2;
      }
      if (tri_flags & GEMM_B_UPPER) {            // op(B)[k,n] = 0 for k > n: live iff k <= c0 + 7
        if (k0 > col_hi) mb = 0;
        else if (k0 + 5 > col_lo) mb = 2;
      }
      mode = (ma == 0 || mb == 0) ? 0 : (ma == 2 ? 2 : (mb == 2 ? 3 : 1));
    }
    if (mode == 1) {
      if (prune == 1) tile_full<T, TA, TB, HAS_W, 1>(acc, sA, sB, sW, a_base, b_base, y2, z2, t);
      else if (prune == 2) tile_full<T, TA, TB, HAS_W, 2>(acc, sA, sB, sW, a_base, b_base, y2, z2, t);
      else tile_full<T, TA, TB, HAS_W, 0>(acc, sA, sB, sW, a_base, b_base, y2, z2, t);
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < T::MT; i++) {
        const int r0 = m0 + 8 * (i * T::WGM + wr);   // rows r0 .. r0 + 7 of this MMA block
        int lo = 0, hi = TBK;
        if (tri_flags & GEMM_A_LOWER) {          // k4-step at k is live iff k <= r0 + 7
          const int u = r0 + 7 - k0;
          hi = u < 0 ? 0 : min(TBK, (u / 4 + 1) * 4);
        }
        if (tri_flags & GEMM_A_UPPER) {          // live iff k + 3 >= r0
          const int v = r0 - k0 - 3;
          lo = v <= 0 ? 0 : ((v + 3) / 4) * 4;
        }
#pragma unroll 1
        for (int kk = lo; kk < hi; kk += 4) {
          const int ak = koff<!TA>(kk, y2, z2), bk = koff<TB>(kk, y2, z2);
          const double af = sA[a_base + i * T::AS + ak];
          double bf[T::NT];
#pragma unroll
          for (int j = 0; j < T::NT; j++) bf[j] = sB[b_base + j * T::BS + bk];
          if (HAS_W) {
            const double w = sW[kk + t];
#pragma unroll
            for (int j = 0; j < T::NT; j++) bf[j] *= w;
          }
#pragma unroll
          for (int j = 0; j < T::NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
        }
      }
    }
    else if (mode == 3) {
#pragma unroll
      for (int j = 0; j < T::NT; j++) {
        const int c0 = n0 + 8 * (j * WGN + wc);      // columns c0 .. c0 + 7 of this MMA block
        int lo = 0, hi = TBK;
        if (tri_flags & GEMM_B_UPPER) {
          const int u = c0 + 7 - k0;
          hi = u < 0 ? 0 : min(TBK, (u / 4 + 1) * 4);
        }
        if (tri_flags & GEMM_B_LOWER) {
          const int v = c0 - k0 - 3;
          lo = v <= 0 ? 0 : ((v + 3) / 4) * 4;
        }
#pragma unroll 1
        for (int kk = lo; kk < hi; kk += 4) {
          const int ak = koff<!TA>(kk, y2, z2), bk = koff<TB>(kk, y2, z2);
          double bf = sB[b_base + j * T::BS + bk];
          if (HAS_W) bf *= sW[kk + t];
          double af[T::MT];
#pragma unroll
          for (int i = 0; i < T::MT; i++) af[i] = sA[a_base + i * T::AS + ak];
#pragma unroll
          for (int i = 0; i < T::MT; i++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);      // this warp's reads of the stage are complete
    if (lane == 0 && warp == kt % (2 * WGN)) {   // re-arm the stage released one k-tile ago (duty rotates over the warps)
      const int L = kt + STAGES - 1;
      if (L < nk) {
        if (L >= STAGES) mbar_wait(&empty[L % STAGES], (unsigned)((L / STAGES - 1) & 1));
        issue(L);
      }
    }
    __syncwarp();
  }

  if (T::CAN_SPLIT && splitk > 1) {
    __syncthreads();                                   // every warp is done with the operand ring (all issued loads were consumed)
    double* red = smem;                                // [MT][NT][2][NTH]
    if (krank != 0) {
#pragma unroll
      for (int i = 0; i < T::MT; i++)
#pragma unroll
        for (int j = 0; j < T::NT; j++) {
          red[((i * T::NT + j) * 2 + 0) * T::NTH + threadIdx.x] = acc[i][j][0];
          red[((i * T::NT + j) * 2 + 1) * T::NTH + threadIdx.x] = acc[i][j][1];
        }
    }
    cluster_sync_all();                                // partial sums visible across the cluster
    if (krank == 0) {
      const unsigned mine = smem_u32(red + threadIdx.x);
      for (int r = 1; r < splitk; r++) {
        const unsigned theirs = mapa_u32(mine, (unsigned)r);
#pragma unroll
        for (int i = 0; i < T::MT; i++)
#pragma unroll
          for (int j = 0; j < T::NT; j++) {
            acc[i][j][0] += ld_dsmem_f64(theirs + (unsigned)(((i * T::NT + j) * 2 + 0) * T::NTH * 8));
            acc[i][j][1] += ld_dsmem_f64(theirs + (unsigned)(((i * T::NT + j) * 2 + 1) * T::NTH * 8));
          }
      }
    }
    cluster_sync_all();                                // the leader has read everything: the other CTAs may retire
    if (krank != 0) return;
  }

  // ---- epilogue (same contract as gemm.cu); lane (g, t) holds C[row(g)][col(2t)], C[row(g)][col(2t + 1)] per block
  double* Cg = p.C + (long long)b * p.sC;
  const double alpha = p.alpha * (p.alpha_vec ? p.alpha_vec[b] : 1.0);
  const double beta = p.beta, gamma = p.gamma * (p.gamma_vec ? p.gamma_vec[b] : 1.0);
  const double* Aux = p.Aux ? p.Aux + (long long)b * p.sAux : nullptr;
  const double* cs = p.colscale ? p.colscale + (long long)b * p.sColscale : nullptr;
  const double* rv = p.rowvec ? p.rowvec + (long long)b * p.sRowvec : nullptr;
  const double* cv = p.colvec ? p.colvec + (long long)b * p.sColvec : nullptr;
  const bool mirror = (p.flags & GEMM_C_MIRROR) && c_lower;
  const bool zero_upper = (p.flags & GEMM_ZERO_UPPER) && c_lower;
  const bool vecC = !TB && ((p.ldc & 1) == 0) && ((((uintptr_t)Cg) & 15) == 0) && !c_lower;
  const int rg = TA ? nu : rho;                                  // row inside the 8-row block
  const int g0 = 2 * t, g1 = 2 * t + 1;                          // MMA columns held by this lane -> rho / nu of them
  const int c0l = TB ? (2 * (g0 & 3) + (g0 >> 2)) : (g0 & 1) + 4 * ((g0 >> 1) & 1) + 2 * (g0 >> 2);
  const int c1l = TB ? (2 * (g1 & 3) + (g1 >> 2)) : c0l + 1;
#pragma unroll
  for (int i = 0; i < T::MT; i++) {
    const int row = m0 + 8 * (i * T::WGM + wr) + rg;
    if (row >= M) continue;
    const double rvv = rv ? rv[row] : 0.0;
#pragma unroll
    for (int j = 0; j < T::NT; j++) {
      const int cb = n0 + 8 * (j * WGN + wc);
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = cb + (e ? c1l : c0l);
        double x = alpha * acc[i][j][e];
        if (c < N) {
          if (Aux) x += gamma * Aux[(long long)row * p.ldaux + c];
          if (cs) x *= cs[c];
          if (rv) x += rvv * cv[c];
          if (beta != 0.0) x += beta * Cg[(long long)row * p.ldc + c];
        }
        v[e] = x;
      }
      if (vecC && cb + c1l < N) {
        *reinterpret_cast<double2*>(Cg + (long long)row * p.ldc + cb + c0l) = make_double2(v[0], v[1]);
      } else {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = cb + (e ? c1l : c0l);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Operand stored [rows, cols] row-major with leading dimension ld (elements), batch stride sX (0 = shared).
//   k_contig: rows = the m/n dimension, cols = k   -> box {16 k, box_rows}, 128-byte swizzle
//   else    : rows = k, cols = the m/n dimension   -> box {8 columns, 16 k}, 64-byte swizzle
// *one_box (m/n-contiguous operands only): the whole tile is one 4-d box {8, 16 k, box_rows / 8, 1} over the view
// (c % 8, k, c / 8, batch) of the operand -- possible when cols % 8 == 0; otherwise one 3-d box per 8-column block.
bool make_map(CUtensorMap* map, const double* base, int rows, int cols, int ld, long long sX, int batch, bool k_contig,
              int box_rows, bool* one_box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if ((((uintptr_t)base) & 15) || (ld & 1) || (sX & 1) || rows < 1 || cols < 1) return false;
  const cuuint64_t nb = sX ? (cuuint64_t)batch : 1;
  *one_box = false;
  static const int allow4d = getenv("GPX_TMA_4D") ? atoi(getenv("GPX_TMA_4D")) : 1;
  if (!k_contig && allow4d && (cols % 8) == 0) {
    cuuint64_t dims[4] = {8, (cuuint64_t)rows, (cuuint64_t)(cols / 8), nb};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 8, 64, sX ? (cuuint64_t)sX * 8 : (cuuint64_t)ld * 8 * (cuuint64_t)rows};
    cuuint32_t box[4] = {8, (cuuint32_t)TBK, (cuuint32_t)(box_rows / 8), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (strides[0] < (1ull << 40) && strides[2] < (1ull << 40) && !(strides[2] & 15) &&
        fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      *one_box = true;
      return true;
    }
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, nb};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8, sX ? (cuuint64_t)sX * 8 : (cuuint64_t)ld * 8 * (cuuint64_t)rows};
  if (strides[0] >= (1ull << 40) || strides[1] >= (1ull << 40) || (strides[1] & 15)) return false;
  cuuint32_t box[3] = {(cuuint32_t)(k_contig ? TBK : 8), (cuuint32_t)(k_contig ? box_rows : TBK), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, k_contig ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BM, int BN, int WGN, bool TA, bool TB, bool HAS_W, int STAGES>
int launch_tma_cfg(const GemmArgs& a, cudaStream_t st) {
  using T = TCfg<BM, BN, WGN, TA, TB, HAS_W, STAGES>;
  alignas(64) CUtensorMap mapA, mapB;
  bool oneA = false, oneB = false;
  // op(A) is M x K: stored [M, K] (k-contiguous) or, with TRANS_A, [K, M]
  if (!make_map(&mapA, a.A, TA ? a.K : a.M, TA ? a.M : a.K, a.lda, a.sA, a.batch, !TA, BM, &oneA)) return 1;
  // op(B) is K x N: stored [K, N] or, with TRANS_B, [N, K] (k-contiguous)
  if (!make_map(&mapB, a.B, TB ? a.N : a.K, TB ? a.K : a.N, a.ldb, a.sB, a.batch, TB, BN, &oneB)) return 1;
  auto kern = gemm_tma_kernel<BM, BN, WGN, TA, TB, HAS_W, STAGES>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  const int n_nt = (a.N + BN - 1) / BN, n_mt = (a.M + BM - 1) / BM;
  const long long blocks = (long long)n_nt * n_mt * a.batch;
  if (blocks > 0x7fffffffLL) return 1;
  const int ob = (oneA ? 1 : 0) | (oneB ? 2 : 0);
  // split-K over a thread-block cluster when the launch has few tiles and a long k-loop (single-window SYRKs)
  static const int splitk_on = getenv("GPX_GEMM_SPLITK") ? atoi(getenv("GPX_GEMM_SPLITK")) : 1;
  int S = 1;
  const int nk_all = (a.K + TBK - 1) / TBK;
  static const int sk_min = getenv("GPX_SPLITK_MINNK") ? atoi(getenv("GPX_SPLITK_MINNK")) : 12;   // tuning knobs
  static const int sk_div = getenv("GPX_SPLITK_DIV") ? atoi(getenv("GPX_SPLITK_DIV")) : 4;
  if (T::CAN_SPLIT && splitk_on && blocks <= 64 && nk_all >= sk_min) {
    S = nk_all / sk_div;
    if (S > 8) S = 8;
    if (S > 296 / (int)blocks) S = 296 / (int)blocks;
    if (S < 2) S = 1;
  }
  if (S > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(blocks * S));
    cfg.blockDim = dim3(T::NTH);
    cfg.dynamicSmemBytes = T::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, a, mapA, mapB, n_nt, n_mt, ob, S) != cudaSuccess) {
      cudaGetLastError();
      S = 1;                                           // (cluster launch refused: plain launch below)
    }
  }
  if (S == 1) kern<<<(unsigned)blocks, T::NTH, T::SMEM, st>>>(a, mapA, mapB, n_nt, n_mt, ob, 1);
  GPX_CHECK_LAUNCH();
  ++g_gemm_tma_launches;
  return GPX_OK;
}

template <int BM, int BN, int WGN, int STAGES>
int launch_tma_trans(const GemmArgs& a, cudaStream_t st) {
  const bool ta = a.flags & GEMM_TRANS_A, tb = a.flags & GEMM_TRANS_B;
  if (a.kweight) {   // k-weights: A diag(w) B^T form only (operands stored [M,K] and [N,K])
    if (ta || !tb) return 1;
    return launch_tma_cfg<BM, BN, WGN, false, true, true, STAGES>(a, st);
  }
  if (!ta && !tb) return launch_tma_cfg<BM, BN, WGN, false, false, false, STAGES>(a, st);
  if (ta && !tb) return launch_tma_cfg<BM, BN, WGN, true, false, false, STAGES>(a, st);
  if (!ta && tb) return launch_tma_cfg<BM, BN, WGN, false, true, false, STAGES>(a, st);
  return launch_tma_cfg<BM, BN, WGN, true, true, false, STAGES>(a, st);
}

}  // namespace

// Returns GPX_OK / GPX_ERR_* when the TMA kernel was launched (or failed to launch), 1 when the operands do not
// qualify (alignment, odd leading dimension, missing driver entry point): the caller then uses the cp.async kernel.
int launch_gemm_tma(const GemmArgs& a, cudaStream_t st) {
  static const int enabled = getenv("GPX_GEMM_TMA") ? atoi(getenv("GPX_GEMM_TMA")) : 1;
  if (!enabled || a.K < 1) return 1;
  if (a.kweight && ((a.K % TBK) || (((uintptr_t)a.kweight) & 15) || (a.sKw & 1))) return 1;
  const int waste80 = (a.M + 79) / 80 * 80 - a.M, waste128 = (a.M + 127) / 128 * 128 - a.M;
  const bool tri = a.flags & (GEMM_A_LOWER | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_B_UPPER);
  const bool syrk = (a.flags & GEMM_C_LOWER) && !tri;
  if (waste128 < waste80) return (a.N <= 64) ? launch_tma_trans<128, 64, 4, 3>(a, st) : launch_tma_trans<128, 128, 4, 3>(a, st);
  if (a.N <= 64) return launch_tma_trans<80, 64, 4, 3>(a, st);
  if (a.N <= 512 || syrk) {   // M x M x M products and lower-only outputs: square tiles that hug the diagonal
    if (((a.N + 79) / 80 * 80 - a.N) <= ((a.N + 63) / 64 * 64 - a.N)) return launch_tma_trans<80, 80, 2, 3>(a, st);
    return launch_tma_trans<80, 64, 4, 3>(a, st);
  }
  static const int big_stages = getenv("GPX_TMA_STAGES") ? atoi(getenv("GPX_TMA_STAGES")) : 4;   // tuning knob
  if (a.flags & GEMM_TRANS_A) return launch_tma_trans<80, 80, 2, 3>(a, st);
  return big_stages == 3 ? launch_tma_trans<80, 64, 2, 3>(a, st) : launch_tma_trans<80, 64, 2, 4>(a, st);
}

}  // namespace gpx

// Hyper-parameter gradient of the Mercer Matern-1/2 spectral-mixture kernel for inducing points that lie ON the sample
// grid (every gpitch caller: uniform 16 kHz windows, z = decimated samples or init_liv maxima, gpitch/init_models.py:9-51;
// demos/scripts/demo-modgp.py:28,40-41 fixes them).  Replaces tf.gradients through MercerMatern12sm.K
// (gpitch/matern12_spectral_mixture.py:102-117) for the M x N cross-covariance.
//
//   K[m,n] = var * exp(-r[m,n]) * c(z_m - x_n),   c(d) = sum_q e_q cos(w_q d),   w_q = fl(2 pi f_q)
//
// With z_m = x[iz_m] the cosine factor depends on (m, n) only through the integer lag iz_m - n, so every sum over the
// M x N elements that the gradient needs collapses onto two lag histograms
//   D0[lag] = sum W,   D1[lag] = sum W s / r,   W = Kbar_eff[m,n] exp(-r[m,n])      (r in the caller's distance mode,
//                                                                                   i.e. GPflow's rounded expansion)
// built in ONE streaming pass over Kbar (~30 FP64 operations per element instead of ~25 + 7 Q), followed by an
// O(lags x Q) tail:  d var = sum D0 c,  d len = var / l sum D1 c,  d e_q = var sum D0 cos(w_q d),
// d f_q = -2 pi var e_q sum D0 d sin(w_q d).   The pass is independent of Q; thread = lag, so the histograms live in
// registers (no atomics): thread `lag` walks the diagonal n = iz_m - lag of Kbar, adjacent threads read adjacent columns.
// The exponential keeps the reference's per-element arithmetic (its rounding noise at absolute time stamps matters at
// 1e-8); the cosine factor is evaluated at the exact lag distance, which moves each weight by <= ulp(t) w_q ~ 1e-9 at
// t = 240 s (the reference's own argument rounding) -- a relative perturbation of the gradient of that size, not
// amplified by any solve because Kbar is given.
#include "builder.cuh"
#include "fastmath.cuh"
#include <cmath>

namespace gpx {
namespace {

#define TWO_PI_L 6.283185307179586

__device__ __forceinline__ double sqdist_ref_l(double m2zt, double zt2, double xt, double xt2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(m2zt, xt), zt2), xt2);
}

// xt[b, p, n] = x[b / divB, n] / lengthscale[b, p]  (IEEE division, as GPflow scales X before the distance)
__global__ void scaled_cols_kernel(const double* __restrict__ pts, int n, int div, const double* __restrict__ hyp, int P,
                                   int HS, double* __restrict__ xt) {
  const int b = blockIdx.z, p = blockIdx.y;
  const double ls = hyp[((long long)b * P + p) * HS + 1];
  const double* x = pts + (long long)(b / div) * n;
  double* o = xt + ((long long)b * P + p) * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) o[i] = x[i] / ls;
}

constexpr int LT = 256;     // lags per CTA (one per thread)
constexpr int RU = 4;       // rows in flight per thread

struct LagArgs {
  const int* iz;        // [ceil(batch / divA), nA]
  const double* delta;  // [ceil(batch / divB)]
  const double* xt;     // [batch, P, nB]
  double* D;            // [batch, P, 2, nlag]
  int nlag;
  int chunks;           // row chunks per (batch entry, component): > 1 for launches too small to fill the GPU; the CTAs of
                        // the chunks form a thread-block cluster and reduce their histograms through distributed shared memory
};

template <int PC>
__global__ void __launch_bounds__(LT) grad_lag_bin_kernel(const KernArgs a, const LagArgs g) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.z, ch = blockIdx.y % g.chunks, p0 = (blockIdx.y / g.chunks) * PC;
  const int M = a.nA, N = a.nB, HS = 2 + 2 * a.Q;
  double* sT = sm;                         // exp table [64]
  double* sZ = sT + 64;                    // zt, zt^2, -2 zt per row: [M][3] (PC == 1) or [PC][3][M]
  const int Mp = (M + RU - 1) / RU * RU;   // table rows padded to the row-group size with finite (zero) entries
  double* sRow = sZ + PC * 3 * Mp;         // [M] epilogue row vector
  double* sRed = sRow + M;                 // [PC][2][LT] this CTA's histograms, read by the cluster leader (chunks > 1)
  int* sIz = reinterpret_cast<int*>(sRed + PC * 2 * LT);   // [M]
  load_exp_table(sT);
  const double* zrow = a.ptsA + (long long)(b / a.divA) * M;
  const int* izrow = g.iz + (long long)(b / a.divA) * M;
  const bool epi = a.epi_col != nullptr;
  for (int i = threadIdx.x; i < Mp; i += LT) {
    if (i < M) {
      sIz[i] = izrow[i];
      sRow[i] = (epi && a.epi_rowv) ? a.epi_rowv[(long long)b * M + i] : 0.0;
    }
    const double z = (i < M) ? zrow[i] : 0.0;
#pragma unroll
    for (int c = 0; c < PC; c++) {
      const int p = min(p0 + c, a.P - 1);
      const double zt = z / a.hyp[((long long)b * a.P + p) * HS + 1];
      const int o = (PC == 1) ? i * 3 : c * 3 * Mp + i, st = (PC == 1) ? 1 : Mp;    // [Mp][3] for one component, else [PC][3][Mp]
      sZ[o] = zt;
      sZ[o + st] = __dmul_rn(zt, zt);
      sZ[o + 2 * st] = -2.0 * zt;
    }
  }
  __syncthreads();
  const int lag = blockIdx.x * LT + threadIdx.x;       // array index; signed lag = lag - (N - 1)
  const int sl = lag - (N - 1);
  const double* Kb = a.K + (long long)b * a.sK;
  const double* ecol = epi ? a.epi_col + (long long)b * N : nullptr;
  const double* ecolv = (epi && a.epi_colv) ? a.epi_colv + (long long)b * N : nullptr;
  const double* xt[PC];
#pragma unroll
  for (int c = 0; c < PC; c++) xt[c] = g.xt + ((long long)b * a.P + min(p0 + c, a.P - 1)) * N;
  double D0[PC], D1[PC];
#pragma unroll
  for (int c = 0; c < PC; c++) D0[c] = D1[c] = 0.0;

  // rows of this CTA: all of them, or one of g.chunks slices (single-window launches: more CTAs, shorter serial walks)
  const int rpc = ((M + g.chunks - 1) / g.chunks + RU - 1) / RU * RU;
  const int mlo = ch * rpc, mhi = min(M, mlo + rpc);
  for (int m0 = mlo; m0 < mhi; m0 += RU) {
    int n[RU];
    bool v[RU];
    bool any = false;
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int m = m0 + u;
      n[u] = (m < mhi) ? sIz[m] - sl : -1;
      v[u] = (m < mhi) && n[u] >= 0 && n[u] < N;
      any |= v[u];
    }
    if (!__any_sync(0xffffffffu, any)) continue;
    double kb[RU];
#pragma unroll
    for (int u = 0; u < RU; u++) kb[u] = v[u] ? Kb[(long long)(m0 + u) * a.ldk + n[u]] : 0.0;
    if (epi) {
#pragma unroll
      for (int u = 0; u < RU; u++)
        if (v[u]) {
          const double ev = ecolv ? ecolv[n[u]] : 0.0;
          kb[u] = fma(a.epi_alpha * ecol[n[u]], kb[u], sRow[m0 + u] * ev);
        }
    }
#pragma unroll
    for (int c = 0; c < PC; c++) {
      double x[RU];
#pragma unroll
      for (int u = 0; u < RU; u++) x[u] = v[u] ? xt[c][n[u]] : 0.0;
#pragma unroll
      for (int u = 0; u < RU; u++) {
        const int zs = (PC == 1) ? 1 : Mp;
        // straight-line code for all RU x PC pairs (a per-pair `continue` made the compiler re-materialise every 64-bit
        // constant inside each pair's region: 83 -> ~45 instructions per pair).  Pairs off the matrix have kb = x = 0 and
        // a finite (padded) table row, so they add exactly 0.
        const double* zr = sZ + ((PC == 1) ? (m0 + u) * 3 : c * 3 * Mp + m0 + u);
        double s;
        if (a.mode == DIST_REFERENCE) s = sqdist_ref_l(zr[2 * zs], zr[zs], x[u], __dmul_rn(x[u], x[u]));
        else { const double d = zr[0] - x[u]; s = d * d; }
        double rinv;
        const double r = sqrt_pos_rinv(s + 1e-12, rinv);
        const double W = kb[u] * exp_neg<PC == 1>(r, sT);      // constant-bank coefficients pay off in the 1-component pass only
        D0[c] += W;
        D1[c] = fma(W, s * rinv, D1[c]);
      }
    }
  }
  if (g.chunks > 1) {        // cluster of the row chunks of this (lag block, component group): the leader sums the histograms
    if (ch != 0) {
#pragma unroll
      for (int c = 0; c < PC; c++) { sRed[(c * 2 + 0) * LT + threadIdx.x] = D0[c]; sRed[(c * 2 + 1) * LT + threadIdx.x] = D1[c]; }
    }
    cluster_sync_all();
    if (ch == 0) {
      const unsigned mine = smem_u32(sRed + threadIdx.x);
      for (int r = 1; r < g.chunks; r++) {
        const unsigned theirs = mapa_u32(mine, (unsigned)r);
#pragma unroll
        for (int c = 0; c < PC; c++) {
          D0[c] += ld_dsmem_f64(theirs + (unsigned)((c * 2 + 0) * LT * 8));
          D1[c] += ld_dsmem_f64(theirs + (unsigned)((c * 2 + 1) * LT * 8));
        }
      }
    }
    cluster_sync_all();
    if (ch != 0) return;
  }
  if (lag < g.nlag) {
#pragma unroll
    for (int c = 0; c < PC; c++) {
      if (p0 + c >= a.P) break;
      double* D = g.D + ((long long)b * a.P + p0 + c) * 2 * g.nlag;
      D[lag] = D0[c];
      D[g.nlag + lag] = D1[c];
    }
  }
}

// One CTA per (component, batch entry): reduce the lag histograms against cos / sin of the lag distance.
constexpr int TQ = 10;
__global__ void __launch_bounds__(256) grad_lag_tail_kernel(const KernArgs a, const LagArgs g) {
  __shared__ double sRed[(2 * TQ + 2) * 8];
  __shared__ double sTot[2];
  const int b = blockIdx.y, p = blockIdx.x;
  const int Q = a.Q, HS = 2 + 2 * Q, N = a.nB;
  const double* h = a.hyp + ((long long)b * a.P + p) * HS;
  double* dh = a.dhyp + ((long long)b * a.P + p) * HS;
  const double var = h[0], ls = h[1];
  const double delta = g.delta[b / a.divB];
  const double* D0 = g.D + ((long long)b * a.P + p) * 2 * g.nlag;
  const double* D1 = D0 + g.nlag;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double tvar = 0.0, tlen = 0.0;
  for (int q0 = 0; q0 < Q; q0 += TQ) {
    // cos / sin of w_q d_l along this thread's lags l = tid, tid + 256, ...: one sincos at the first lag, then a rotation by
    // the fixed angle w_q (256 delta) per step (4 FMAs instead of a ~45-operation sincos; 256 delta is exact, the rotation
    // adds <= ~1 ulp of absolute error per step -- 3e-14 after the 256 steps of a 32k-sample window).
    double eq[TQ], ae[TQ], af[TQ], cs[TQ], sn[TQ], cD[TQ], sD[TQ];
    const double dstep = 256.0 * delta;
    const double dfirst = (double)((int)threadIdx.x - (N - 1)) * delta;
#pragma unroll
    for (int q = 0; q < TQ; q++) {
      const bool ok = q0 + q < Q;
      eq[q] = ok ? h[2 + q0 + q] : 0.0;
      const double wq = ok ? __dmul_rn(TWO_PI_L, h[2 + Q + q0 + q]) : 0.0;
      ae[q] = af[q] = 0.0;
      sincos(wq * dfirst, &sn[q], &cs[q]);
      sincos(wq * dstep, &sD[q], &cD[q]);
    }
    double avar = 0.0, alen = 0.0;
    for (int l = threadIdx.x; l < g.nlag; l += 256) {
      const double d0 = D0[l], d1 = D1[l];
      if (d0 != 0.0 || d1 != 0.0) {
        const double d = (double)(l - (N - 1)) * delta;
        double k = 0.0;
#pragma unroll
        for (int q = 0; q < TQ; q++) {
          k = fma(eq[q], cs[q], k);
          ae[q] = fma(d0, cs[q], ae[q]);
          af[q] = fma(d0 * d, sn[q], af[q]);
        }
        avar = fma(d0, k, avar);
        alen = fma(d1, k, alen);
      }
#pragma unroll
      for (int q = 0; q < TQ; q++) {
        const double c = cs[q], sq = sn[q];
        cs[q] = fma(c, cD[q], -sq * sD[q]);
        sn[q] = fma(sq, cD[q], c * sD[q]);
      }
    }
    // block reduction of 2 TQ + 2 values
    double vals[2 * TQ + 2];
#pragma unroll
    for (int q = 0; q < TQ; q++) { vals[q] = ae[q]; vals[TQ + q] = af[q]; }
    vals[2 * TQ] = avar; vals[2 * TQ + 1] = alen;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2 * TQ + 2; j++) {
      const double s = warp_sum(vals[j]);
      if (lane == 0) sRed[j * 8 + w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 2 * TQ + 2) {
      double s = 0.0;
      for (int k = 0; k < 8; k++) s += sRed[threadIdx.x * 8 + k];
      const int j = threadIdx.x;
      if (j < TQ) {
        if (q0 + j < Q) dh[2 + q0 + j] = a.need_ef ? var * s : 0.0;                               // dK/de_q = var E cos
      } else if (j < 2 * TQ) {
        const int q = j - TQ;
        if (q0 + q < Q) dh[2 + Q + q0 + q] = a.need_ef ? -var * TWO_PI_L * h[2 + q0 + q] * s : 0.0;   // dK/df_q
      } else {
        sTot[j - 2 * TQ] = s;                // var / len sums of this chunk of partials
      }
    }
    __syncthreads();
    tvar += sTot[0];
    tlen += sTot[1];
  }
  if (threadIdx.x == 0) {
    dh[0] = tvar;
    dh[1] = var * tlen / ls;
  }
}

}  // namespace

// a: the arguments of launch_kernel_grad (kind must be KIND_MERCER_M12; features are not needed).  iz / delta / scratch as
// described in include/gpitch_b200.h (gpx_kernel_grad_lag).  dhyp is overwritten (no atomics).
int launch_kernel_grad_lag(const KernArgs& a, const int* iz, const double* delta, double* work, int nlag, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0 || a.nB <= 0) return GPX_OK;
  if (a.kind != KIND_MERCER_M12 || a.P < 1 || a.Q < 1 || !a.dhyp || !iz || !delta || !work || nlag < a.nB) return GPX_ERR_ARG;
  if (a.P > 65535) return GPX_ERR_ARG;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  LagArgs g;
  g.iz = iz; g.delta = delta; g.nlag = nlag;
  double* xt = work;                                              // [batch, P, nB]
  g.xt = xt;
  g.D = work + (long long)a.batch * a.P * a.nB;                   // [batch, P, 2, nlag]
  {
    // launches that would not fill the GPU (a single window): up to 8 row chunks per (lag block, component group), their
    // CTAs in one thread-block cluster
    const int pcs = (a.P >= 4) ? 4 : 1;
    const long long ctas = (long long)((nlag + LT - 1) / LT) * ((a.P + pcs - 1) / pcs) * a.batch;
    const int want = a.nA / 32;
    g.chunks = (ctas >= 148 || want < 2) ? 1 : (want > 8 ? 8 : want);
  }
  const int HS = 2 + 2 * a.Q;
  for (int b0 = 0; b0 < a.batch; b0 += 65535) {                   // grid.z limit
    const int nb = a.batch - b0 < 65535 ? a.batch - b0 : 65535;
    KernArgs s = a;
    LagArgs gs = g;
    s.batch = nb;
    s.hyp += (long long)b0 * a.P * HS; s.dhyp += (long long)b0 * a.P * HS;
    s.K += (long long)b0 * a.sK;
    if (b0 % a.divA || b0 % a.divB) return GPX_ERR_ARG;           // (65535-entry slices must not split a window)
    s.ptsA += (long long)(b0 / a.divA) * a.nA; s.ptsB += (long long)(b0 / a.divB) * a.nB;
    gs.iz += (long long)(b0 / a.divA) * a.nA; gs.delta += b0 / a.divB;
    if (a.epi_col) s.epi_col += (long long)b0 * a.nB;
    if (a.epi_rowv) s.epi_rowv += (long long)b0 * a.nA;
    if (a.epi_colv) s.epi_colv += (long long)b0 * a.nB;
    gs.xt += (long long)b0 * a.P * a.nB; gs.D += (long long)b0 * a.P * 2 * nlag;
    double* xts = xt + (long long)b0 * a.P * a.nB;
    scaled_cols_kernel<<<dim3((a.nB + 255) / 256, a.P, nb), 256, 0, st>>>(s.ptsB, a.nB, a.divB, s.hyp, a.P, HS, xts);
    GPX_CHECK_LAUNCH();
    const int pc = (a.P >= 4) ? 4 : 1;
    const size_t smem = ((size_t)64 + (size_t)pc * 3 * ((a.nA + RU - 1) / RU * RU) + a.nA + (size_t)pc * 2 * LT) * sizeof(double) +
                        (size_t)a.nA * sizeof(int);
    if (smem > 200 * 1024) return GPX_ERR_ARG;
    dim3 grid((nlag + LT - 1) / LT, ((a.P + pc - 1) / pc) * g.chunks, nb);
    {
      void (*kern)(const KernArgs, const LagArgs) = (pc == 4) ? grad_lag_bin_kernel<4> : grad_lag_bin_kernel<1>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(LT);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1;
      attr[0].val.clusterDim.y = (unsigned)gs.chunks;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = gs.chunks > 1 ? 1 : 0;
      if (cudaLaunchKernelEx(&cfg, kern, s, gs) != cudaSuccess) return GPX_ERR_LAUNCH;
    }
    GPX_CHECK_LAUNCH();
    grad_lag_tail_kernel<<<dim3(a.P, nb), 256, 0, st>>>(s, gs);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}

}  // namespace gpx

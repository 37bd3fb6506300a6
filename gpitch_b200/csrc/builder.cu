#include "builder.cuh"
#include "fastmath.cuh"
#include <cmath>
#include <type_traits>

namespace gpx {

#define TWO_PI 6.283185307179586  /* Python: 2 * np.pi */

// ---------------------------------------------------------------------------------------------------------
// Mercer features  phi[q] = sqrt(e_q) cos(fl(fl(2 pi f_q) x)),  phi[Q+q] = sqrt(e_q) sin(...)   (absolute time,
// same operation order as MercerMatern12sm.phi_features, matern12_spectral_mixture.py:123-133).
// ---------------------------------------------------------------------------------------------------------
__global__ void features_kernel(const double* __restrict__ pts, int n, int div, const double* __restrict__ hyp,
                                int P, int Q, double* __restrict__ feat) {
  const int b = blockIdx.z, p = blockIdx.y;
  const int KP = (2 * Q + 3) / 4 * 4, HS = 2 + 2 * Q;
  const double* h = hyp + ((long long)b * P + p) * HS;
  const double* x = pts + (long long)(b / div) * n;
  double* f = feat + ((long long)b * P + p) * KP * (long long)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double xi = x[i];
    for (int q = 0; q < Q; q++) {
      const double se = sqrt(h[2 + q]);
      const double w = __dmul_rn(TWO_PI, h[2 + Q + q]);
      double s, c;
      sincos(__dmul_rn(w, xi), &s, &c);
      f[(long long)q * n + i] = se * c;
      f[(long long)(Q + q) * n + i] = se * s;
    }
    for (int q = 2 * Q; q < KP; q++) f[(long long)q * n + i] = 0.0;
  }
}

int launch_features(const double* pts, int n, int div, const double* hyp, int P, int Q, double* feat, int batch,
                    cudaStream_t st) {
  if (batch <= 0 || n <= 0 || Q <= 0) return GPX_OK;
  dim3 grid((n + 255) / 256, P, batch);
  features_kernel<<<grid, 256, 0, st>>>(pts, n, div, hyp, P, Q, feat);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// scaled squared distance, reference order: ((-2 * (zt * xt)) + zt^2) + xt^2 with every op rounded separately
// (GPflow Stationary.square_dist; TF executes each op as its own kernel, so no FMA contraction ever happens).
__device__ __forceinline__ double sqdist_ref(double m2zt, double zt2, double xt, double xt2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(m2zt, xt), zt2), xt2);
}

// ---------------------------------------------------------------------------------------------------------
// Builder: K[b] = sum_p k_p(ptsA, ptsB).  One CTA owns a 128-column strip and walks row tiles of 32 inducing
// points; 8 warps (2 x 4), warp tile 16 x 32.  The rank-2Q feature contraction runs on DMMA.
//
// exp(-r) without a per-element sqrt / exp ("separable" tiles).  With d = |z/l - x/l| (an exact fp64 difference)
// and r = sqrt(s + 1e-12) the reference distance (s by expansion in `reference` mode, d^2 in `stable` mode):
//     exp(-r) = exp(-d) * exp(-(r - d)),    r - d = eps = q/2 * (1 - q h / 4) + O(d eta^3),
//     q = (s + 1e-12 - d^2) * h,  h = 1/d,  eta = q h   (series of d sqrt(1 + eta) - d)
// exp(-d) factorises over a tile that lies on one side of the diagonal (all x >= all z or the reverse):
// exp(-d) = u_m v_n with u_m = exp(-|c - z_m/l|), v_n = exp(-|x_n/l - c|), c = the strip edge facing the rows, so
// u_m (and the variance) are folded into the row features once per row tile, v_n is a per-column table, and the
// per-element work is ~17 FP64 operations + the DMMA share instead of ~29 + DMMA.  eps <= 1e-5 carries the whole
// difference between the reference's rounded expansion and the exact distance, so the result equals
// variance * exp(-r_reference) * k to ~1e-15 relative.  Elements with |eta| > 2^-15 (coincident / nearly coincident
// points) and tiles that straddle the diagonal take the exact sqrt_pos / exp_neg path.
// ---------------------------------------------------------------------------------------------------------
constexpr int BBM = 32, BBN = 128, BTHREADS = 256, BMT = 2;
constexpr int B_LDA = BBM + 4, B_LDB = BBN + 4;
constexpr double SEP_SPAN_MAX = 600.0;   // largest |c - z/l| for which 1/u_m stays finite

__device__ __forceinline__ double rcp_approx(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

template <int KIND, int MODE, bool P1>
__global__ void __launch_bounds__(BTHREADS, 2) build_kernel(const KernArgs a, const int rt_per_cta) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[8];
  __shared__ int s_sep;   // 0: exact path for the whole tile, +1: all x >= all z, -1: all x <= all z
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * BBN;
  const int Q = a.Q, HS = 2 + 2 * Q;
  constexpr bool MERCER = KIND == KIND_MERCER_M12;
  constexpr double CEXP = (KIND == KIND_MATERN32) ? 1.7320508075688772 : 1.0;   // exponent scale: exp(-CEXP r)
  const int KP = MERCER ? (2 * Q + 3) / 4 * 4 : 0;
  double* sFA = sm;                    // [KP][B_LDA]  row features (x variance x u_m in separable tiles)
  double* sFB = sFA + KP * B_LDA;      // [KP][B_LDB]
  double* sZ = sFB + KP * B_LDB;       // per row: z, zt, zt^2, -2 zt, u, 1/u            [6][BBM]
  double* sX = sZ + 6 * BBM;           // per col: x, xt, xt^2, v+ (x >= z side), v- (x <= z side)  [5][BBN]
  double* sH = sX + 5 * BBN;           // hypers of the current component [HS]
  double* sT = sH + HS;                // 2^(j/64) table for exp_neg
  load_exp_table(sT);

  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * (8 * BMT), wn0 = (warp & 3) * 32;
  double* Kg = a.K + (long long)b * a.sK;
  const bool vec = ((a.ldk & 1) == 0) && ((((uintptr_t)Kg) & 15) == 0);
  const int nA = a.nA, nB = a.nB;

  const int n_rt = (nA + BBM - 1) / BBM;
  const int rt0 = blockIdx.y * rt_per_cta, rt1 = min(n_rt, rt0 + rt_per_cta);
  bool x_staged = false;
  double xt_min = 0.0, xt_max = 0.0;   // scaled-coordinate range of the valid columns of this strip

  for (int rt = rt0; rt < rt1; rt++) {
    const int m0 = rt * BBM;
    double tot[P1 ? 1 : BMT][P1 ? 1 : 4][2];
    const int Pn = P1 ? 1 : a.P;

    for (int p = 0; p < Pn; p++) {
      __syncthreads();  // previous component / row tile fully consumed
      const double* h = a.hyp + ((long long)b * a.P + p) * HS;
      for (int i = threadIdx.x; i < HS; i += BTHREADS) sH[i] = h[i];
      const double ls = h[1], var = h[0];
      const bool stage_x = !(P1 && x_staged);
      if (stage_x) {
        // column side: scaled inputs, their range (warp 0..3 own 32 columns each), v tables
        double lo = 1e300, hi = -1e300;
        if (threadIdx.x < BBN) {
          const int c = n0 + threadIdx.x;
          const double x = (c < nB) ? xrow[c] : 0.0, xt = x / ls;
          sX[threadIdx.x] = x; sX[BBN + threadIdx.x] = xt; sX[2 * BBN + threadIdx.x] = __dmul_rn(xt, xt);
          if (c < nB) { lo = xt; hi = xt; }
        }
        if (threadIdx.x < BBN) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
          }
          if (lane == 0) { s_red[warp] = lo; s_red[4 + warp] = hi; }
        }
        __syncthreads();
        xt_min = fmin(fmin(s_red[0], s_red[1]), fmin(s_red[2], s_red[3]));
        xt_max = fmax(fmax(s_red[4], s_red[5]), fmax(s_red[6], s_red[7]));
        if (KIND != KIND_DIFF_M12 && threadIdx.x < BBN) {
          const double xt = sX[BBN + threadIdx.x];
          sX[3 * BBN + threadIdx.x] = exp_neg(CEXP * fmax(xt - xt_min, 0.0), sT);   // rows below the strip
          sX[4 * BBN + threadIdx.x] = exp_neg(CEXP * fmax(xt_max - xt, 0.0), sT);   // rows above the strip
        }
        if (MERCER) {
          const double* fb = a.featB + ((long long)b * a.P + p) * KP * (long long)nB;
          for (int idx = threadIdx.x; idx < KP * BBN; idx += BTHREADS) {
            int k = idx / BBN, i = idx - k * BBN;
            sFB[k * B_LDB + i] = (n0 + i < nB) ? fb[(long long)k * nB + n0 + i] : 0.0;
          }
        }
        x_staged = true;
      }
      // row side: scaled inputs + range (warp 0) -> tile classification -> u, 1/u
      if (warp == 0) {
        const int r = m0 + lane;
        const bool rv = r < nA;
        const double z = rv ? zrow[r] : 0.0, zt = z / ls;
        sZ[lane] = z; sZ[BBM + lane] = zt; sZ[2 * BBM + lane] = __dmul_rn(zt, zt); sZ[3 * BBM + lane] = -2.0 * zt;
        double lo = rv ? zt : 1e300, hi = rv ? zt : -1e300;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
          hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        int sep = 0;
        double u = 1.0, ru = 1.0;
        if (KIND != KIND_DIFF_M12) {
          if (xt_min >= hi && CEXP * (xt_min - lo) < SEP_SPAN_MAX) sep = 1;          // all x >= all z
          else if (xt_max <= lo && CEXP * (hi - xt_max) < SEP_SPAN_MAX) sep = -1;    // all x <= all z
          if (sep != 0 && rv) {
            const double e = CEXP * (sep > 0 ? xt_min - zt : zt - xt_max);           // >= 0
            u = exp_neg(e, sT);
            ru = exp(e);
          }
        }
        sZ[4 * BBM + lane] = u; sZ[5 * BBM + lane] = ru;
        if (lane == 0) s_sep = sep;
      }
      __syncthreads();
      const int sep = s_sep;
      if (MERCER) {
        const double* fa = a.featA + ((long long)b * a.P + p) * KP * (long long)nA;
        for (int idx = threadIdx.x; idx < KP * BBM; idx += BTHREADS) {
          int k = idx / BBM, i = idx - k * BBM;
          const double f = (m0 + i < nA) ? fa[(long long)k * nA + m0 + i] : 0.0;
          sFA[k * B_LDA + i] = f * (var * sZ[4 * BBM + i]);      // variance and u_m folded into the row features
        }
        __syncthreads();
      }

      double acc[BMT][4][2];
      if (MERCER) {
#pragma unroll
        for (int i = 0; i < BMT; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < KP; kk += 4) {
          double af[BMT], bf[4];
#pragma unroll
          for (int i = 0; i < BMT; i++) af[i] = sFA[(kk + t) * B_LDA + wm0 + i * 8 + g];
#pragma unroll
          for (int j = 0; j < 4; j++) bf[j] = sFB[(kk + t) * B_LDB + wn0 + j * 8 + g];
#pragma unroll
          for (int i = 0; i < BMT; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
      const double* sV = sX + (sep > 0 ? 3 : 4) * BBN;
#pragma unroll
      for (int i = 0; i < BMT; i++) {
        const int rl = wm0 + i * 8 + g;
        const double z = sZ[rl], zt = sZ[BBM + rl], zt2 = sZ[2 * BBM + rl], m2zt = sZ[3 * BBM + rl];
        const double um = sZ[4 * BBM + rl], rum = sZ[5 * BBM + rl];
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cl = wn0 + j * 8 + 2 * t + e;
            double kv;
            if (KIND == KIND_DIFF_M12) {
              // Matern12sm.K: r = |z - x + 1e-12|; var * exp(-r/l) * sum_q e_q cos(2 pi f_q r)   (:47-56)
              const double r = fabs(__dadd_rn(__dadd_rn(z, -sX[cl]), 1e-12));
              double k = 0.0;
              for (int q = 0; q < Q; q++) {
                const double ph = __dmul_rn(__dmul_rn(TWO_PI, sH[2 + Q + q]), r);
                const double term = sH[2 + q] * cos(ph);
                k = (q == 0) ? term : k + term;
              }
              if (a.kind == KIND_DIFF_M32) {      // Matern32sm.K (gpitch/kernels.py:230-242): (1 + r1) exp(-r1), r1 = sqrt(3) r / l
                const double r1 = 1.7320508075688772 * (r / sH[1]);
                kv = ((var * (1.0 + r1)) * exp(-r1)) * k;
              } else {
                kv = (var * exp(-(r / sH[1]))) * k;
              }
            } else {
              const double xt = sX[BBN + cl];
              const double d = fabs(zt - xt);
              double s;
              if (MODE == DIST_REFERENCE) s = sqdist_ref(m2zt, zt2, xt, sX[2 * BBN + cl]);
              else s = d * d;
              const double sp = s + 1e-12;
              bool done = false;
              if (sep != 0) {
                const double hh = rcp_approx(d);
                const double q = fma(-d, d, sp) * hh;            // 2 eps_0
                const double w = q * hh;                          // eta
                if (fabs(w) < 3.0517578125e-05) {                 // 2^-15 (NaN / inf from d == 0 fail the test)
                  const double eps2 = q * fma(w, -0.25, 1.0);     // 2 (r - d)
                  if (MERCER) {
                    // exp(-eps) = 1 - eps + eps^2/2,  eps = eps2 / 2
                    const double corr = fma(eps2, fma(eps2, 0.125, -0.5), 1.0);
                    kv = acc[i][j][e] * (sV[cl] * corr);
                  } else {
                    // var (1 + c r) exp(-c r),  c = sqrt(3), r = d + eps,  exp(-c eps) = 1 - c eps + (c eps)^2 / 2
                    const double ce = (0.5 * CEXP) * eps2;
                    const double corr = fma(ce, fma(ce, 0.5, -1.0), 1.0);
                    kv = (var * um) * (sV[cl] * corr) * (1.0 + fma(CEXP, d, ce));
                  }
                  done = true;
                }
              }
              if (!done) {
                const double r = sqrt_pos(sp);
                if (MERCER) kv = (exp_neg(r, sT) * rum) * acc[i][j][e];
                else { const double s3r = CEXP * r; kv = (var * (1.0 + s3r)) * exp_neg(s3r, sT); }
              }
            }
            if (P1) acc[i][j][e] = kv;
            else tot[i][j][e] = (p == 0) ? kv : tot[i][j][e] + kv;
          }
        }
      }
      if (P1) {
#pragma unroll
        for (int i = 0; i < BMT; i++) {
          const int row = m0 + wm0 + i * 8 + g;
          if (row >= nA) continue;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int col = n0 + wn0 + j * 8 + 2 * t;
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            if (a.jitter != 0.0) { if (row == col) v0 += a.jitter; if (row == col + 1) v1 += a.jitter; }
            double* dst = Kg + (long long)row * a.ldk + col;
            if (vec && col + 1 < nB) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            else { if (col < nB) dst[0] = v0; if (col + 1 < nB) dst[1] = v1; }
          }
        }
      }
    }
    if (!P1) {
#pragma unroll
      for (int i = 0; i < BMT; i++) {
        const int row = m0 + wm0 + i * 8 + g;
        if (row >= nA) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int col = n0 + wn0 + j * 8 + 2 * t;
          double v0 = tot[P1 ? 0 : i][P1 ? 0 : j][0], v1 = tot[P1 ? 0 : i][P1 ? 0 : j][1];
          if (a.jitter != 0.0) { if (row == col) v0 += a.jitter; if (row == col + 1) v1 += a.jitter; }
          double* dst = Kg + (long long)row * a.ldk + col;
          if (vec && col + 1 < nB) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
          else { if (col < nB) dst[0] = v0; if (col + 1 < nB) dst[1] = v1; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-component builder (P == 1: every Pdgp latent GP, SGPR with one pitch).  Same arithmetic as above, but
// the whole row side of the CTA (scaled inducing inputs, per-tile side-of-diagonal classification, u_m tables) is
// prepared once per CTA, and the Mercer row features of tile rt+1 stream in with cp.async while tile rt is
// contracted, so the only per-tile synchronisation is one barrier.
// ---------------------------------------------------------------------------------------------------------
constexpr int RT_MAX = 16;   // row tiles per CTA (shared-memory row tables are sized for RT_MAX * 32 rows)

template <int KIND, int MODE>
__global__ void __launch_bounds__(BTHREADS, 3) build_kernel_p1(const KernArgs a, const int rt_per_cta) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[8];
  __shared__ int s_sep[RT_MAX];   // per row tile: 0 exact path, +1 all x >= all z, -1 all x <= all z
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * BBN;
  const int Q = a.Q, HS = 2 + 2 * Q;
  constexpr bool MERCER = KIND == KIND_MERCER_M12;
  constexpr double CEXP = (KIND == KIND_MATERN32) ? 1.7320508075688772 : 1.0;
  const int KP = MERCER ? (2 * Q + 3) / 4 * 4 : 0;
  constexpr int RROWS = RT_MAX * BBM;
  double* sFA = sm;                        // [2][KP][B_LDA]  double-buffered row features
  double* sFB = sFA + 2 * KP * B_LDA;      // [KP][B_LDB]
  double* sZ = sFB + KP * B_LDB;           // per row: zt, zt^2, -2 zt, var*u, var/u      [5][RROWS]
  double* sX = sZ + 5 * RROWS;             // per col: xt, xt^2, v+, v-                    [4][BBN]
  double* sT = sX + 4 * BBN;               // 2^(j/64) table
  load_exp_table(sT);

  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const double* h = a.hyp + (long long)b * HS;
  const double var = h[0], ls = h[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * (8 * BMT), wn0 = (warp & 3) * 32;
  double* Kg = a.K + (long long)b * a.sK;
  const bool vec = ((a.ldk & 1) == 0) && ((((uintptr_t)Kg) & 15) == 0);
  const int nA = a.nA, nB = a.nB;
  const int n_rt = (nA + BBM - 1) / BBM;
  const int rt0 = blockIdx.y * rt_per_cta, rt1 = min(n_rt, rt0 + rt_per_cta);
  const int ntile = rt1 - rt0;
  const double* fa = MERCER ? a.featA + (long long)b * KP * (long long)nA : nullptr;

  auto prefetch_features = [&](int tile, int buf) {     // row features of tile -> sFA[buf]  (8-byte cp.async)
    if (MERCER) {
      const int m0 = (rt0 + tile) * BBM;
      double* dst = sFA + buf * KP * B_LDA;
      for (int idx = threadIdx.x; idx < KP * BBM; idx += BTHREADS) {
        const int k = idx / BBM, i = idx - k * BBM;
        const bool v = m0 + i < nA;
        cp_async8(dst + k * B_LDA + i, fa + (long long)k * nA + (v ? m0 + i : 0), v ? 8 : 0);
      }
    }
  };
  if (ntile > 0) prefetch_features(0, 0);
  cp_async_commit();

  // ---- column side (once per CTA)
  {
    double lo = 1e300, hi = -1e300;
    if (threadIdx.x < BBN) {
      const int c = n0 + threadIdx.x;
      const double xt = ((c < nB) ? xrow[c] : 0.0) / ls;
      sX[threadIdx.x] = xt; sX[BBN + threadIdx.x] = __dmul_rn(xt, xt);
      if (c < nB) { lo = xt; hi = xt; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == 0) { s_red[warp] = lo; s_red[4 + warp] = hi; }
    }
    if (MERCER) {
      const double* fb = a.featB + (long long)b * KP * (long long)nB;
      for (int idx = threadIdx.x; idx < KP * BBN; idx += BTHREADS) {
        int k = idx / BBN, i = idx - k * BBN;
        sFB[k * B_LDB + i] = (n0 + i < nB) ? fb[(long long)k * nB + n0 + i] : 0.0;
      }
    }
  }
  __syncthreads();
  const double xt_min = fmin(fmin(s_red[0], s_red[1]), fmin(s_red[2], s_red[3]));
  const double xt_max = fmax(fmax(s_red[4], s_red[5]), fmax(s_red[6], s_red[7]));
  if (threadIdx.x < BBN) {
    const double xt = sX[threadIdx.x];
    sX[2 * BBN + threadIdx.x] = exp_neg(CEXP * fmax(xt - xt_min, 0.0), sT);   // tiles with all z <= the strip
    sX[3 * BBN + threadIdx.x] = exp_neg(CEXP * fmax(xt_max - xt, 0.0), sT);   // tiles with all z >= the strip
  }
  // ---- row side (once per CTA): warp w prepares row tiles w, w + 8, ...
  for (int tile = warp; tile < ntile; tile += BTHREADS / 32) {
    const int r = (rt0 + tile) * BBM + lane;
    const bool rv = r < nA;
    const double zt = (rv ? zrow[r] : 0.0) / ls;
    double lo = rv ? zt : 1e300, hi = rv ? zt : -1e300;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    int sep = 0;
    if (xt_min >= hi && CEXP * (xt_min - lo) < SEP_SPAN_MAX) sep = 1;
    else if (xt_max <= lo && CEXP * (hi - xt_max) < SEP_SPAN_MAX) sep = -1;
    double vu = var, vru = var;
    if (sep != 0 && rv) {
      const double e = CEXP * (sep > 0 ? xt_min - zt : zt - xt_max);   // >= 0
      vu = var * exp_neg(e, sT);
      vru = var * exp(e);
    }
    const int rl = tile * BBM + lane;
    sZ[rl] = zt; sZ[RROWS + rl] = __dmul_rn(zt, zt); sZ[2 * RROWS + rl] = -2.0 * zt;
    sZ[3 * RROWS + rl] = vu; sZ[4 * RROWS + rl] = vru;
    if (lane == 0) s_sep[tile] = sep;
  }

  for (int tile = 0; tile < ntile; tile++) {
    const int m0 = (rt0 + tile) * BBM;
    const int buf = tile & 1;
    cp_async_wait<0>();                    // this tile's features have landed (own copies) ...
    __syncthreads();                       // ... and everybody else's; tile - 1 is fully consumed (sFA[buf ^ 1] free)
    if (tile + 1 < ntile) prefetch_features(tile + 1, buf ^ 1);
    cp_async_commit();
    const int sep = s_sep[tile];
    const double* sV = sX + (sep > 0 ? 2 : 3) * BBN;
    const double* cA = sFA + buf * KP * B_LDA;
    // The warp's 16 x 32 tile is produced in two 16 x 16 halves to keep the live register set small (3 CTAs / SM).
#pragma unroll 1
    for (int jh = 0; jh < 2; jh++) {
      const int wnh = wn0 + jh * 16;
      double acc[BMT][2][2];
      if (MERCER) {
#pragma unroll
        for (int i = 0; i < BMT; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < KP; kk += 4) {
          double af[BMT], bf[2];
#pragma unroll
          for (int i = 0; i < BMT; i++) af[i] = cA[(kk + t) * B_LDA + wm0 + i * 8 + g];
#pragma unroll
          for (int j = 0; j < 2; j++) bf[j] = sFB[(kk + t) * B_LDB + wnh + j * 8 + g];
#pragma unroll
          for (int i = 0; i < BMT; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
      // The side-of-diagonal class is uniform over the tile: two straight-line code paths, no per-element branch.
      // In separable tiles the (rare) elements that fail the series test are patched afterwards.
      unsigned bad = 0;
#pragma unroll
      for (int i = 0; i < BMT; i++) {
        const int rl = tile * BBM + wm0 + i * 8 + g;
        const double zt = sZ[rl], zt2 = sZ[RROWS + rl], m2zt = sZ[2 * RROWS + rl];
        const double vu = sZ[3 * RROWS + rl];
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cl = wnh + j * 8 + 2 * t + e;
            const double xt = sX[cl];
            const double d = fabs(zt - xt);
            double s;
            if (MODE == DIST_REFERENCE) s = sqdist_ref(m2zt, zt2, xt, sX[BBN + cl]);
            else s = d * d;
            const double sp = s + 1e-12;
            double kv;
            bool isbad = false;
            if (sep != 0) {
              const double hh = rcp_approx(d);
              const double q = fma(-d, d, sp) * hh;            // 2 eps_0
              const double w = q * hh;                          // eta
              isbad = !(fabs(w) < 3.0517578125e-05);            // 2^-15; NaN / inf (d == 0) are bad too
              if (isbad) bad |= 1u << ((i * 2 + j) * 2 + e);
              const double eps2 = q * fma(w, -0.25, 1.0);       // 2 (r - d)
              if (MERCER) {
                const double corr = fma(eps2, fma(eps2, 0.125, -0.5), 1.0);
                kv = acc[i][j][e] * (vu * (sV[cl] * corr));
              } else {
                const double ce = (0.5 * CEXP) * eps2;
                const double corr = fma(ce, fma(ce, 0.5, -1.0), 1.0);
                kv = vu * (sV[cl] * corr) * (1.0 + fma(CEXP, d, ce));
              }
            } else {
              const double r = sqrt_pos(sp);
              if (MERCER) kv = (exp_neg(r, sT) * var) * acc[i][j][e];
              else { const double s3r = CEXP * r; kv = (var * (1.0 + s3r)) * exp_neg(s3r, sT); }
            }
            if (!(MERCER && isbad)) acc[i][j][e] = kv;          // Mercer: keep the raw contraction for the patch
          }
        }
      }
      if (bad) {   // exact path for the flagged elements
#pragma unroll
        for (int i = 0; i < BMT; i++) {
          const int rl = tile * BBM + wm0 + i * 8 + g;
          const double zt = sZ[rl], zt2 = sZ[RROWS + rl], m2zt = sZ[2 * RROWS + rl];
#pragma unroll
          for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              if (!((bad >> ((i * 2 + j) * 2 + e)) & 1u)) continue;
              const int cl = wnh + j * 8 + 2 * t + e;
              const double xt = sX[cl];
              double s;
              if (MODE == DIST_REFERENCE) s = sqdist_ref(m2zt, zt2, xt, sX[BBN + cl]);
              else { const double d = zt - xt; s = d * d; }
              const double r = sqrt_pos(s + 1e-12);
              if (MERCER) acc[i][j][e] = (exp_neg(r, sT) * var) * acc[i][j][e];
              else { const double s3r = CEXP * r; acc[i][j][e] = (var * (1.0 + s3r)) * exp_neg(s3r, sT); }
            }
        }
      }
#pragma unroll
      for (int i = 0; i < BMT; i++) {
        const int row = m0 + wm0 + i * 8 + g;
        if (row >= nA) continue;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int col = n0 + wnh + j * 8 + 2 * t;
          double v0 = acc[i][j][0], v1 = acc[i][j][1];
          if (a.jitter != 0.0) { if (row == col) v0 += a.jitter; if (row == col + 1) v1 += a.jitter; }
          double* dst = Kg + (long long)row * a.ldk + col;
          if (vec && col + 1 < nB) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
          else { if (col < nB) dst[0] = v0; if (col + 1 < nB) dst[1] = v1; }
        }
      }
    }
  }
  cp_async_wait<0>();
}

template <int KIND, int MODE>
static int launch_build_p1(const KernArgs& a, cudaStream_t st) {
  const int KP = (KIND == KIND_MERCER_M12) ? feat_rows(a.Q) : 0;
  size_t smem = ((size_t)KP * (2 * B_LDA + B_LDB) + 5 * RT_MAX * BBM + 4 * BBN + 64) * sizeof(double);
  if (smem > 200 * 1024) return GPX_ERR_ARG;
  auto kern = build_kernel_p1<KIND, MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n_rt = (a.nA + BBM - 1) / BBM, strips = (a.nB + BBN - 1) / BBN;
  long long tiles = (long long)n_rt * strips * a.batch;
  int rt_per = (int)(tiles / 900);
  rt_per = rt_per < 1 ? 1 : (rt_per > n_rt ? n_rt : rt_per);
  if (rt_per > RT_MAX) rt_per = RT_MAX;
  dim3 grid(strips, (n_rt + rt_per - 1) / rt_per, a.batch);
  kern<<<grid, BTHREADS, smem, st>>>(a, rt_per);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Multi-component builder (GPflow Add of P pitch kernels: SGPRSS in gpitch/transcription.py:245, P up to 88).
// One CTA owns one 32 x 128 output tile and sums the P components in registers.  Per component the Mercer feature
// tiles of BOTH sides stream in with cp.async (double buffered across p) while the previous component is
// evaluated; the scaled coordinates, the separable-exp tables u_m / v_n and the side-of-diagonal class are rebuilt
// per component (the lengthscale changes) by 160 threads between two barriers.  Arithmetic as in build_kernel_p1.
// ---------------------------------------------------------------------------------------------------------
template <int KIND, int MODE>
__global__ void __launch_bounds__(BTHREADS, 3) build_kernel_sum(const KernArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * BBN, m0 = blockIdx.y * BBM;
  const int Q = a.Q, HS = 2 + 2 * Q, P = a.P;
  constexpr bool MERCER = KIND == KIND_MERCER_M12;
  constexpr double CEXP = (KIND == KIND_MATERN32) ? 1.7320508075688772 : 1.0;
  const int KP = MERCER ? (2 * Q + 3) / 4 * 4 : 0;
  const int FSZ = KP * (B_LDA + B_LDB);
  double* sF = sm;                         // [2][ KP x B_LDA  |  KP x B_LDB ]  double-buffered feature tiles
  double* sZ = sF + 2 * FSZ;               // per row: raw z, zt, zt^2, -2 zt, var*u, var/u      [6][BBM]
  double* sX = sZ + 6 * BBM;               // per col: raw x, xt, xt^2, v, 1/v                   [5][BBN]
  double* sT = sX + 5 * BBN;               // 2^(j/64) table
  load_exp_table(sT);

  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * (8 * BMT), wn0 = (warp & 3) * 32;
  const int nA = a.nA, nB = a.nB;

  auto prefetch_features = [&](int p, int buf) {
    if (MERCER) {
      const double* fa = a.featA + ((long long)b * P + p) * KP * (long long)nA;
      const double* fb = a.featB + ((long long)b * P + p) * KP * (long long)nB;
      double* dA = sF + buf * FSZ;
      double* dB = dA + KP * B_LDA;
      for (int idx = threadIdx.x; idx < KP * BBM; idx += BTHREADS) {
        const int k = idx / BBM, i = idx - k * BBM;
        const bool v = m0 + i < nA;
        cp_async8(dA + k * B_LDA + i, fa + (long long)k * nA + (v ? m0 + i : 0), v ? 8 : 0);
      }
      for (int idx = threadIdx.x; idx < KP * BBN; idx += BTHREADS) {
        const int k = idx / BBN, i = idx - k * BBN;
        const bool v = n0 + i < nB;
        cp_async8(dB + k * B_LDB + i, fb + (long long)k * nB + (v ? n0 + i : 0), v ? 8 : 0);
      }
    }
  };
  prefetch_features(0, 0);
  cp_async_commit();

  // raw coordinates and their ranges (geometry does not depend on the component); warps 0..3 = columns, warp 4 = rows
  __shared__ double s_lo[5], s_hi[5];
  {
    double lo = 1e300, hi = -1e300;
    if (threadIdx.x < BBN) {
      const int c = n0 + threadIdx.x;
      const double x = (c < nB) ? xrow[c] : 0.0;
      sX[threadIdx.x] = x;
      if (c < nB) { lo = x; hi = x; }
    } else if (threadIdx.x < BBN + BBM) {
      const int r = m0 + threadIdx.x - BBN;
      const double z = (r < nA) ? zrow[r] : 0.0;
      sZ[threadIdx.x - BBN] = z;
      if (r < nA) { lo = z; hi = z; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0 && warp < 5) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
  }
  const double x_min = fmin(fmin(s_lo[0], s_lo[1]), fmin(s_lo[2], s_lo[3]));
  const double x_max = fmax(fmax(s_hi[0], s_hi[1]), fmax(s_hi[2], s_hi[3]));
  const double z_min = s_lo[4], z_max = s_hi[4];
  const int sep = (x_min >= z_max) ? 1 : ((x_max <= z_min) ? -1 : 0);   // dividing by l > 0 keeps the order

  double tot[BMT][4][2];
#pragma unroll
  for (int i = 0; i < BMT; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) tot[i][j][0] = tot[i][j][1] = 0.0;

  for (int p = 0; p < P; p++) {
    const int buf = p & 1;
    const double* h = a.hyp + ((long long)b * P + p) * HS;
    const double var = h[0], ls = h[1];
    cp_async_wait<0>();
    __syncthreads();                       // features(p) landed everywhere; component p - 1 fully consumed
    if (p + 1 < P) prefetch_features(p + 1, buf ^ 1);
    cp_async_commit();
    // per-component tables.  Separable tiles (all x on one side of all z): exp(-|xt - zt|) = u_m v_n relative to the tile
    // edge.  Mixed tiles (the band around the diagonal, ~1/5 of the tiles at M = 200, N = 2001) get TWO-SIDED tables relative
    // to the left column edge c:  x >= z: exp(zt - c) exp(-(xt - c)),  x < z: exp(-(zt - c)) exp(xt - c)  -- the element picks
    // its side, so these tiles run the same short series as the separable ones instead of a sqrt + exp per element
    // (that path remains for coincident / nearly coincident points and for tiles too wide for the lengthscale).
    const double xt_edge = (sep > 0 ? x_min : (sep < 0 ? x_max : x_min)) / ls;
    const bool mixok = sep == 0 && CEXP * (fmax(x_max, z_max) - fmin(x_min, z_min)) < 300.0 * ls;
    if (threadIdx.x < BBN) {
      const double xt = sX[threadIdx.x] / ls;
      sX[BBN + threadIdx.x] = xt; sX[2 * BBN + threadIdx.x] = __dmul_rn(xt, xt);
      double v = 1.0, vi = 1.0;
      if (sep != 0) v = exp_neg(CEXP * fmax(sep > 0 ? xt - xt_edge : xt_edge - xt, 0.0), sT);
      else if (mixok) { v = exp_neg(CEXP * fmax(xt - xt_edge, 0.0), sT); vi = 1.0 / v; }
      sX[3 * BBN + threadIdx.x] = v; sX[4 * BBN + threadIdx.x] = vi;
    } else if (threadIdx.x < BBN + BBM) {
      const int i = threadIdx.x - BBN;
      const double zt = sZ[i] / ls;
      sZ[BBM + i] = zt; sZ[2 * BBM + i] = __dmul_rn(zt, zt); sZ[3 * BBM + i] = -2.0 * zt;
      double u = 1.0, ui = 1.0;
      if (sep != 0) u = exp_neg(CEXP * fmax(sep > 0 ? xt_edge - zt : zt - xt_edge, 0.0), sT);
      else if (mixok) {
        const double aa = CEXP * (zt - xt_edge);                   // either sign: u = exp(aa), ui = exp(-aa)
        const double en = exp_neg(fabs(aa), sT);
        u = aa >= 0.0 ? 1.0 / en : en;
        ui = aa >= 0.0 ? en : 1.0 / en;
      }
      sZ[4 * BBM + i] = var * u; sZ[5 * BBM + i] = var * ui;
    }
    __syncthreads();

    // The warp's 16 x 32 tile is evaluated in two 16 x 16 halves (fully unrolled: `tot` keeps static register indices) so
    // that only half of the contraction accumulators are live at a time: 3 CTAs per SM instead of 2 -- this kernel is bound
    // by dependent-latency stalls of its FP64 chains (ncu: stall_wait 34 %), i.e. by the number of resident warps.
    // (Tried and reverted: double-buffered tables written for component p + 1 while p is evaluated, one barrier per component
    // instead of two -- 4.92 -> 5.27 ms at P = 88: the five warps that rebuild the tables arrive late at the next barrier.)
    // The tile class is branched on OUTSIDE the element loops (three straight-line instances of the element code): with the
    // branch inside, the compiler re-materialised the 64-bit constants in every element's region.
    const double* cA = sF + buf * FSZ;
    const double* cB = cA + KP * B_LDA;
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      double acc[BMT][2][2];
      if (MERCER) {
#pragma unroll
        for (int i = 0; i < BMT; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < KP; kk += 4) {
          double af[BMT], bf[2];
#pragma unroll
          for (int i = 0; i < BMT; i++) af[i] = cA[(kk + t) * B_LDA + wm0 + i * 8 + g];
#pragma unroll
          for (int j = 0; j < 2; j++) bf[j] = cB[(kk + t) * B_LDB + wn0 + (jh * 2 + j) * 8 + g];
#pragma unroll
          for (int i = 0; i < BMT; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
      auto elements = [&](auto em_) {
        constexpr int EM = decltype(em_)::value;       // 0: exact, 1: separable tile, 2: mixed tile, two-sided tables
#pragma unroll
        for (int i = 0; i < BMT; i++) {
          const int rl = wm0 + i * 8 + g;
          const double zt = sZ[BBM + rl], zt2 = sZ[2 * BBM + rl], m2zt = sZ[3 * BBM + rl], vu = sZ[4 * BBM + rl];
          const double vui = (EM == 2) ? sZ[5 * BBM + rl] : 0.0;
#pragma unroll
          for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int cl = wn0 + (jh * 2 + j) * 8 + 2 * t + e;
              const double xt = sX[BBN + cl];
              const double d = fabs(zt - xt);
              double s;
              if (MODE == DIST_REFERENCE) s = sqdist_ref(m2zt, zt2, xt, sX[2 * BBN + cl]);
              else s = d * d;
              const double sp = s + 1e-12;
              double kv;
              bool isbad = false;
              if (EM != 0) {
                const double hh = rcp_approx(d);
                const double q = fma(-d, d, sp) * hh;
                const double w = q * hh;
                isbad = !(fabs(w) < 3.0517578125e-05);
                const double eps2 = q * fma(w, -0.25, 1.0);
                double uv;
                if (EM == 2) {
                  const bool xge = xt >= zt;
                  uv = (xge ? vu : vui) * (xge ? sX[3 * BBN + cl] : sX[4 * BBN + cl]);
                } else {
                  uv = vu * sX[3 * BBN + cl];
                }
                if (MERCER) {
                  const double corr = fma(eps2, fma(eps2, 0.125, -0.5), 1.0);
                  kv = acc[i][j][e] * (uv * corr);
                } else {
                  const double ce = (0.5 * CEXP) * eps2;
                  const double corr = fma(ce, fma(ce, 0.5, -1.0), 1.0);
                  kv = (uv * corr) * (1.0 + fma(CEXP, d, ce));
                }
              }
              if (EM == 0 || isbad) {                          // exact path (rare in the tabulated tiles: coincident points)
                const double r = sqrt_pos(sp);
                if (MERCER) kv = (exp_neg(r, sT) * var) * acc[i][j][e];
                else { const double s3r = CEXP * r; kv = (var * (1.0 + s3r)) * exp_neg(s3r, sT); }
              }
              tot[i][jh * 2 + j][e] = (p == 0) ? kv : tot[i][jh * 2 + j][e] + kv;
            }
          }
        }
      };
      if (sep != 0) elements(std::integral_constant<int, 1>{});
      else if (mixok) elements(std::integral_constant<int, 2>{});
      else elements(std::integral_constant<int, 0>{});
    }
  }
  cp_async_wait<0>();

  double* Kg = a.K + (long long)b * a.sK;
  const bool vec = ((a.ldk & 1) == 0) && ((((uintptr_t)Kg) & 15) == 0);
#pragma unroll
  for (int i = 0; i < BMT; i++) {
    const int row = m0 + wm0 + i * 8 + g;
    if (row >= nA) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = n0 + wn0 + j * 8 + 2 * t;
      double v0 = tot[i][j][0], v1 = tot[i][j][1];
      if (a.jitter != 0.0) { if (row == col) v0 += a.jitter; if (row == col + 1) v1 += a.jitter; }
      double* dst = Kg + (long long)row * a.ldk + col;
      if (vec && col + 1 < nB) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      else { if (col < nB) dst[0] = v0; if (col + 1 < nB) dst[1] = v1; }
    }
  }
}

template <int KIND, int MODE>
static int launch_build_sum(const KernArgs& a, cudaStream_t st) {
  const int KP = (KIND == KIND_MERCER_M12) ? feat_rows(a.Q) : 0;
  size_t smem = ((size_t)2 * KP * (B_LDA + B_LDB) + 6 * BBM + 5 * BBN + 64) * sizeof(double);
  if (smem > 200 * 1024) return GPX_ERR_ARG;
  auto kern = build_kernel_sum<KIND, MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.nB + BBN - 1) / BBN, (a.nA + BBM - 1) / BBM, a.batch);
  kern<<<grid, BTHREADS, smem, st>>>(a);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

template <int KIND, int MODE, bool P1>
static int launch_build_cfg(const KernArgs& a, cudaStream_t st) {
  const int KP = (KIND == KIND_MERCER_M12) ? feat_rows(a.Q) : 0;
  size_t smem = ((size_t)KP * (B_LDA + B_LDB) + 6 * BBM + 5 * BBN + 2 + 2 * a.Q + 64) * sizeof(double);
  if (smem > 200 * 1024) return GPX_ERR_ARG;
  auto kern = build_kernel<KIND, MODE, P1>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // row tiles per CTA: as many as possible while keeping >= ~6 CTAs per SM in flight
  const int n_rt = (a.nA + BBM - 1) / BBM, strips = (a.nB + BBN - 1) / BBN;
  long long tiles = (long long)n_rt * strips * a.batch;
  int rt_per = (int)(tiles / 900);
  rt_per = rt_per < 1 ? 1 : (rt_per > n_rt ? n_rt : rt_per);
  dim3 grid(strips, (n_rt + rt_per - 1) / rt_per, a.batch);
  kern<<<grid, BTHREADS, smem, st>>>(a, rt_per);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

template <int KIND>
static int launch_build_kind(const KernArgs& a, cudaStream_t st) {
  const bool p1 = a.P == 1;
  if (p1 && KIND != KIND_DIFF_M12)
    return a.mode == DIST_REFERENCE ? launch_build_p1<(KIND == KIND_DIFF_M12 ? KIND_MATERN32 : KIND), DIST_REFERENCE>(a, st)
                                    : launch_build_p1<(KIND == KIND_DIFF_M12 ? KIND_MATERN32 : KIND), DIST_STABLE>(a, st);
  if (KIND == KIND_DIFF_M12)
    return p1 ? launch_build_cfg<KIND, DIST_REFERENCE, true>(a, st) : launch_build_cfg<KIND, DIST_REFERENCE, false>(a, st);
  constexpr int K2 = (KIND == KIND_DIFF_M12) ? KIND_MATERN32 : KIND;   // (never instantiated for the difference form)
  return a.mode == DIST_REFERENCE ? launch_build_sum<K2, DIST_REFERENCE>(a, st) : launch_build_sum<K2, DIST_STABLE>(a, st);
}

int launch_kernel_build(const KernArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0 || a.nB <= 0) return GPX_OK;
  if (a.batch > 65535 || a.P < 1) return GPX_ERR_ARG;
  if (a.kind == KIND_MERCER_M12 && (!a.featA || !a.featB || a.Q < 1)) return GPX_ERR_ARG;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  if (a.kind == KIND_MERCER_M12) return launch_build_kind<KIND_MERCER_M12>(a, st);
  if (a.kind == KIND_MATERN32) return launch_build_kind<KIND_MATERN32>(a, st);
  return launch_build_kind<KIND_DIFF_M12>(a, st);
}

// ---------------------------------------------------------------------------------------------------------
// Hyper-parameter gradient: dhyp[b,p,:] += sum_{m,n} Kbar[m,n] dK_p[m,n]/dtheta   (SURVEY Appendix B.1).
// K is never re-read: every term is re-evaluated from the points / features, Kbar is read exactly once.
// One thread per column, GBM rows per CTA; rows are processed four at a time with the four Kbar loads issued
// up front (the loop is otherwise one global-load latency per row); per-row quantities (z/l, (z/l)^2, features)
// are staged in shared memory; register accumulators -> block reduce -> one atomic per value.
// ---------------------------------------------------------------------------------------------------------
constexpr int GBM = 40, GTHREADS = 256, GROWS = 4;

// Sum `n` per-thread values over the CTA with ONE barrier: warp shuffles, per-warp partials in shared memory, then
// thread j < n adds the warp partials of value j and hands the total to `sink(j, total)`.  scratch >= n * 8 doubles.
template <int NV, class Sink>
__device__ __forceinline__ void block_sum_many(const double (&v)[NV], int n, double* scratch, Sink sink) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();                         // scratch free (previous use consumed)
#pragma unroll
  for (int j = 0; j < NV; j++) {
    if (j < n) {
      const double s = warp_sum(v[j]);
      if (lane == 0) scratch[j * 8 + w] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double s = 0.0;
    for (int k = 0; k < nw; k++) s += scratch[threadIdx.x * 8 + k];
    sink(threadIdx.x, s);
  }
}

template <int GQ>
struct GradAcc {
  double e[GQ > 0 ? GQ : 1], f[GQ > 0 ? GQ : 1];
};

// DZ: also accumulate the gradient w.r.t. the row points (trainable inducing inputs) into a.dpts -- same distance /
// exponential / feature work, plus one w-scaled feature read per partial and a warp reduction per row.
template <bool NEED_EF, int GQ, bool DZ>
__global__ void __launch_bounds__(GTHREADS, 2) grad_kernel(const KernArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[32];
  __shared__ double sRed[(2 * (GQ > 0 ? GQ : 1) + 2) * 8];
  __shared__ double sRowv[GBM];
  __shared__ double sDz[GBM];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * GBM, c = blockIdx.x * GTHREADS + threadIdx.x;
  const int Q = a.Q, HS = 2 + 2 * Q;
  const int KP = (a.kind == KIND_MERCER_M12) ? (2 * Q + 3) / 4 * 4 : 0;
  double* sT = sm;                  // exp table [64]
  double* sZ = sT + 64;             // per row: z, z/l, (z/l)^2, -2 z/l   [4][GBM]
  constexpr int GQ1s = GQ > 0 ? GQ : 1;
  const int QP = (GQ > 0) ? (Q + GQ1s - 1) / GQ1s * GQ1s : Q;   // partials padded to whole register chunks
  double* sFA = sZ + 4 * GBM;       // [GBM][2 QP] row features, (cos_q, sin_q) pairs, zero padded
  double* sK = sFA + GBM * 2 * QP;  // [GBM][GTHREADS] Kbar tile: read from HBM once, reused by every component
  double* sFW = sK + GBM * GTHREADS; // DZ: [GBM][2 QP] row features scaled by w_q = 2 pi f_q
  if (DZ && threadIdx.x < GBM) sDz[threadIdx.x] = 0.0;
  load_exp_table(sT);
  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const double* Kb = a.K + (long long)b * a.sK;
  const bool colv = c < a.nB;
  const int cc = colv ? c : 0;
  const double x = colv ? xrow[c] : 0.0;
  const int rows = min(GBM, a.nA - m0);
  // fused epilogue on the adjoint: Kbar_eff = epi_c * Kbar + rowv[m] * epi_v   (identity when no epilogue is given)
  const bool epi = a.epi_col != nullptr;
  const double epi_c = (epi && colv) ? a.epi_alpha * a.epi_col[(long long)b * a.nB + c] : 0.0;
  const double epi_v = (epi && colv && a.epi_colv) ? a.epi_colv[(long long)b * a.nB + c] : 0.0;
  if (epi && threadIdx.x < GBM)
    sRowv[threadIdx.x] = (a.epi_rowv && threadIdx.x < rows) ? a.epi_rowv[(long long)b * a.nA + m0 + threadIdx.x] : 0.0;
  // each thread streams its own Kbar column into shared memory (coalesced across the CTA); it is the only reader of
  // that column, so no barrier is needed -- just the thread's own cp.async completion
  for (int i = 0; i < GBM; i++) {
    const bool v = colv && i < rows;
    cp_async8(sK + i * GTHREADS + threadIdx.x, Kb + (long long)(v ? m0 + i : m0) * a.ldk + cc, v ? 8 : 0);
  }
  cp_async_commit();

  for (int p = 0; p < a.P; p++) {
    const double* h = a.hyp + ((long long)b * a.P + p) * HS;
    double* dh = a.dhyp + ((long long)b * a.P + p) * HS;
    const double var = h[0], ls = h[1];
    __syncthreads();
    for (int i = threadIdx.x; i < GBM; i += GTHREADS) {
      const double z = (i < rows) ? zrow[m0 + i] : 0.0, zt = z / ls;
      sZ[i] = z; sZ[GBM + i] = zt; sZ[2 * GBM + i] = __dmul_rn(zt, zt); sZ[3 * GBM + i] = -2.0 * zt;
    }
    if (a.kind == KIND_MERCER_M12) {
      const double* fa = a.featA + ((long long)b * a.P + p) * KP * (long long)a.nA;
      for (int idx = threadIdx.x; idx < GBM * 2 * QP; idx += GTHREADS) {
        int k = idx / GBM, i = idx - k * GBM;                  // k over [0, 2 QP): cos block then sin block
        const int q = k < QP ? k : k - QP;                     // (cos_q, sin_q) pairs adjacent -> one 16-byte read
        const int ksrc = k < QP ? q : Q + q;
        const double fv = (i < rows && q < Q) ? fa[(long long)ksrc * a.nA + m0 + i] : 0.0;
        sFA[i * 2 * QP + 2 * q + (k < QP ? 0 : 1)] = fv;
        if (DZ) sFW[i * 2 * QP + 2 * q + (k < QP ? 0 : 1)] = (q < Q) ? __dmul_rn(TWO_PI, h[2 + Q + q]) * fv : 0.0;
      }
    }
    __syncthreads();
    cp_async_wait<0>();
    const double xt = x / ls, xt2 = __dmul_rn(xt, xt);
    double a_var = 0.0, a_len = 0.0;

    if (a.kind == KIND_DIFF_M12 || a.kind == KIND_DIFF_M32) {
      // r = |z - x + 1e-12|, K = var g(r) sum_q e_q cos(w_q r) with g = exp(-r/l), dK/dl = K r / l^2, or (Matern32sm)
      // g = (1 + r1) exp(-r1), r1 = sqrt(3) r / l, dK/dl = var k r1^2 exp(-r1) / l   (not on the named path: plain
      // libm trig per element)
      const bool m32 = a.kind == KIND_DIFF_M32;
      for (int q0 = 0; q0 < Q || q0 == 0; q0 += (GQ > 0 ? GQ : 1)) {
        GradAcc<GQ> A;
#pragma unroll
        for (int q = 0; q < (GQ > 0 ? GQ : 1); q++) A.e[q] = A.f[q] = 0.0;
        if (colv)
          for (int i = 0; i < rows; i++) {
            const double r = fabs(__dadd_rn(__dadd_rn(sZ[i], -x), 1e-12));
            double kbe = sK[i * GTHREADS + threadIdx.x];
            if (epi) kbe = fma(epi_c, kbe, sRowv[i] * epi_v);
            const double r1 = m32 ? 1.7320508075688772 * (r / ls) : r / ls, er = exp(-r1);
            const double W = kbe * (m32 ? (1.0 + r1) * er : er);
            double k = 0.0;
            for (int q = 0; q < Q; q++) {
              double sn, cs;
              sincos(__dmul_rn(__dmul_rn(TWO_PI, h[2 + Q + q]), r), &sn, &cs);
              k += h[2 + q] * cs;
              if (NEED_EF && q >= q0 && q < q0 + GQ) { A.e[q - q0] += W * cs; A.f[q - q0] += W * r * sn; }
            }
            if (q0 == 0) { a_var += W * k; a_len += m32 ? kbe * k * r1 * r1 * er * ls : W * k * r; }
          }
        if (NEED_EF)
          for (int q = 0; q < GQ && q0 + q < Q; q++) {
            double se = block_sum<false>(A.e[q], red), sf = block_sum<false>(A.f[q], red);
            if (threadIdx.x == 0) {
              atomicAdd(dh + 2 + q0 + q, var * se);
              atomicAdd(dh + 2 + Q + q0 + q, -var * h[2 + q0 + q] * TWO_PI * sf);
            }
          }
        if (!NEED_EF) break;
      }
      a_var = block_sum<false>(a_var, red);
      a_len = block_sum<false>(a_len, red);
      if (threadIdx.x == 0) { atomicAdd(dh + 0, a_var); atomicAdd(dh + 1, var * a_len / (ls * ls)); }
      continue;
    }

    // ---- Stationary kinds (Mercer Matern-1/2 SM, Matern-3/2): distance, exp and weights per element
    const double* fb = (a.kind == KIND_MERCER_M12) ? a.featB + ((long long)b * a.P + p) * KP * (long long)a.nB : nullptr;
    const bool mercer = a.kind == KIND_MERCER_M12;
    constexpr int GQ1 = GQ > 0 ? GQ : 1;
    const int nchunk = (mercer && GQ > 0) ? (Q + GQ1 - 1) / GQ1 : 1;   // var / len sums are linear in the partials too
    for (int ch = 0; ch < nchunk; ch++) {
      const int q0 = ch * (GQ > 0 ? GQ : 1);
      double xc[GQ > 0 ? GQ : 1], xs[GQ > 0 ? GQ : 1];
      GradAcc<GQ> A;
#pragma unroll
      for (int q = 0; q < (GQ > 0 ? GQ : 1); q++) {
        const bool v = mercer && GQ > 0 && (q0 + q < Q);
        xc[q] = v ? fb[(long long)(q0 + q) * a.nB + cc] : 0.0;
        xs[q] = v ? fb[(long long)(Q + q0 + q) * a.nB + cc] : 0.0;
        A.e[q] = A.f[q] = 0.0;
      }
      for (int i0 = 0; i0 < rows; i0 += GROWS) {
        double kb[GROWS], dzv[DZ ? GROWS : 1];
#pragma unroll
        for (int u = 0; u < GROWS; u++) kb[u] = sK[(i0 + u) * GTHREADS + threadIdx.x];   // zero-filled beyond `rows`
        if (epi) {
#pragma unroll
          for (int u = 0; u < GROWS; u++) kb[u] = fma(epi_c, kb[u], sRowv[i0 + u] * epi_v);    // sRowv = 0 beyond `rows`
        }
#pragma unroll
        for (int u = 0; u < GROWS; u++) {
          const int i = i0 + u;                                  // rows beyond `rows` carry kb = 0 -> no effect
          const double z = sZ[i], zt = sZ[GBM + i];
          if (DZ) dzv[u] = 0.0;
          double s;
          if (a.mode == DIST_REFERENCE) s = sqdist_ref(sZ[3 * GBM + i], sZ[2 * GBM + i], xt, xt2);
          else { const double d = zt - xt; s = d * d; }
          double rinv;
          const double r = sqrt_pos_rinv(s + 1e-12, rinv);
          if (!mercer) {                                         // Matern-3/2: dK/dvar = K/var, dK/dl = 3 var E s / l
            const double s3r = 1.7320508075688772 * r, E = exp_neg(s3r, sT);
            a_var += kb[u] * (1.0 + s3r) * E;
            a_len += kb[u] * E * s;
            if (DZ) dzv[u] = -3.0 * var * kb[u] * E * (zt - xt) / ls;   // dK/dz = -3 var exp(-sqrt(3) r) d~ / l
            continue;
          }
          const double W = kb[u] * exp_neg(r, sT);
          const double* fz = sFA + i * 2 * QP;
          if (GQ > 0) {                                           // column features live in registers
            const double Wd = W * (z - x);
            double k = 0.0, sw = 0.0;
#pragma unroll
            for (int q = 0; q < GQ; q++) {                        // padded partials have zero features: no test
              const double2 zz = *reinterpret_cast<const double2*>(fz + 2 * (q0 + q));
              const double zc = zz.x, zs = zz.y;
              const double cq = fma(zc, xc[q], zs * xs[q]);       // e_q cos(w_q (z - x))
              k += cq;
              if (NEED_EF) {                                      // energies / frequencies trainable
                const double sq = fma(zs, xc[q], -zc * xs[q]);    // e_q sin(w_q (z - x))
                A.e[q] = fma(W, cq, A.e[q]);
                A.f[q] = fma(Wd, sq, A.f[q]);
              }
              if (DZ) {                                           // sum_q w_q e_q sin(w_q (z - x)) from w-scaled features
                const double2 ww = *reinterpret_cast<const double2*>(sFW + i * 2 * QP + 2 * (q0 + q));
                sw = fma(ww.y, xc[q], fma(-ww.x, xs[q], sw));
              }
            }
            const double Wk = W * k;
            a_var += Wk;
            a_len = fma(Wk, s * rinv, a_len);
            if (DZ) dzv[u] = -var * fma(Wk, (zt - xt) * rinv / ls, W * sw);   // dK/dz = -var E [k d~/(l r) + sum_q w_q e_q sin]
          } else {                                               // energies / frequencies fixed: only k is needed
            double k = 0.0;
            for (int q = 0; q < Q; q++)
              k += fz[2 * q] * fb[(long long)q * a.nB + cc] + fz[2 * q + 1] * fb[(long long)(Q + q) * a.nB + cc];
            a_var += W * k;
            a_len += W * k * (s * rinv);
          }
        }
        if (DZ) {                                               // per-row sums over this CTA's columns
#pragma unroll
          for (int u = 0; u < GROWS; u++) {
            const double t = warp_sum(dzv[u]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sDz[i0 + u], t);
          }
        }
      }
      {   // one barrier-pair per component: [e_0..e_GQ-1 | f_0..f_GQ-1 | var | len]
        constexpr int NV = 2 * GQ1 + 2;
        double vals[NV];
#pragma unroll
        for (int q = 0; q < GQ1; q++) { vals[q] = A.e[q]; vals[GQ1 + q] = A.f[q]; }
        vals[2 * GQ1] = a_var; vals[2 * GQ1 + 1] = a_len;
        const bool ef = mercer && NEED_EF && GQ > 0;
        block_sum_many<NV>(vals, NV, sRed, [&](int j, double tot) {
          if (j < GQ1) {
            if (ef && q0 + j < Q) {
              const double eq = h[2 + q0 + j];
              atomicAdd(dh + 2 + q0 + j, eq > 0.0 ? var * tot / eq : 0.0);          // dK/de_q = var E cos
            }
          } else if (j < 2 * GQ1) {
            const int q = j - GQ1;
            if (ef && q0 + q < Q) atomicAdd(dh + 2 + Q + q0 + q, -var * TWO_PI * tot);   // dK/df_q = -var E e_q 2 pi d sin
          } else {                                               // every chunk of partials contributes
            if (j == 2 * GQ1) atomicAdd(dh + 0, tot);
            else atomicAdd(dh + 1, (mercer ? var : 3.0 * var) * tot / ls);
          }
        });
        a_var = 0.0; a_len = 0.0;
      }
    }
  }
  if (DZ) {                                                     // row-point gradient of this CTA's column block
    __syncthreads();
    if (threadIdx.x < rows) atomicAdd(a.dpts + (long long)b * a.nA + m0 + threadIdx.x, sDz[threadIdx.x]);
  }
}

template <bool NEED_EF, int GQ, bool DZ = false>
static int launch_grad_cfg(const KernArgs& a, cudaStream_t st) {
  const int gq1 = GQ > 0 ? GQ : 1;
  const int QP = (GQ > 0) ? (a.Q + gq1 - 1) / gq1 * gq1 : a.Q;
  size_t smem = ((size_t)64 + 4 * GBM + (size_t)GBM * 2 * QP * (DZ ? 2 : 1) + (size_t)GBM * GTHREADS) * sizeof(double);
  cudaFuncSetAttribute(grad_kernel<NEED_EF, GQ, DZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.nB + GTHREADS - 1) / GTHREADS, (a.nA + GBM - 1) / GBM, a.batch);
  grad_kernel<NEED_EF, GQ, DZ><<<grid, GTHREADS, smem, st>>>(a);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

int launch_kernel_grad(const KernArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0 || a.nB <= 0) return GPX_OK;
  if (a.batch > 65535 || a.P < 1 || !a.dhyp) return GPX_ERR_ARG;
  if (a.kind == KIND_MERCER_M12 && (!a.featA || !a.featB || a.Q < 1)) return GPX_ERR_ARG;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  if (a.dpts) {   // fused row-point gradient (trainable inducing inputs): stationary kinds only
    if (a.kind == KIND_MATERN32) return launch_grad_cfg<false, 0, true>(a, st);
    if (a.kind != KIND_MERCER_M12) return GPX_ERR_ARG;
    if (a.Q <= 4) return launch_grad_cfg<true, 4, true>(a, st);
    if (a.Q <= 6) return launch_grad_cfg<true, 6, true>(a, st);
    return launch_grad_cfg<true, 10, true>(a, st);
  }
  if (a.kind == KIND_MATERN32 || (a.kind != KIND_MERCER_M12 && !a.need_ef)) return launch_grad_cfg<false, 0>(a, st);
  if (a.need_ef) {
    if (a.Q <= 4) return launch_grad_cfg<true, 4>(a, st);
    if (a.Q <= 6) return launch_grad_cfg<true, 6>(a, st);
    return launch_grad_cfg<true, 10>(a, st);        // Q <= 10 in one pass; larger Q in chunks of 10 partials
  }
  if (a.Q <= 4) return launch_grad_cfg<false, 4>(a, st);      // energies / frequencies fixed: same register-resident
  if (a.Q <= 6) return launch_grad_cfg<false, 6>(a, st);      // features, without their accumulators
  return launch_grad_cfg<false, 10>(a, st);
}

// ---------------------------------------------------------------------------------------------------------
// Gradient w.r.t. the ROW points (inducing inputs):  dpts[b, m] = sum_p sum_n Kbar[b,m,n] d k_p(z_m, x_n) / d z_m.
// Only needed when a model trains its inducing inputs (gpitch/pdgp.py:80-85 makes za, zc Params; the demos fix
// them), so this is a plain streaming kernel: ZR rows per CTA, threads stride over the columns, the per-row sums
// stay in registers and are block-reduced once.  With r = sqrt(s + 1e-12), s = ((z - x)/l)^2, d~ = (z - x)/l:
//   Mercer Matern-1/2 SM:  dK/dz = -var E [ k d~ / (l r) + sum_q w_q e_q sin(w_q (z - x)) ],  E = exp(-r)
//   Matern-3/2          :  dK/dz = -3 var exp(-sqrt(3) r) d~ / l
// ---------------------------------------------------------------------------------------------------------
constexpr int ZR = 8, ZTHREADS = 256;

__global__ void __launch_bounds__(ZTHREADS) grad_points_kernel(const KernArgs a, double* __restrict__ dpts) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double sRed[ZR * 8];
  __shared__ double sRowv[ZR];
  const int b = blockIdx.y, m0 = blockIdx.x * ZR;
  const int Q = a.Q, HS = 2 + 2 * Q;
  const bool mercer = a.kind == KIND_MERCER_M12;
  const int KP = mercer ? (2 * Q + 3) / 4 * 4 : 0;
  double* sT = sm;                 // exp table [64]
  double* sZ = sT + 64;            // per row: z/l, (z/l)^2, -2 z/l   [3][ZR]
  double* sW = sZ + 3 * ZR;        // angular frequencies [Q]
  double* sFA = sW + Q;            // [ZR][2Q] row features (cos block, sin block)
  load_exp_table(sT);
  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const double* Kb = a.K + (long long)b * a.sK;
  const int rows = min(ZR, a.nA - m0);
  double acc[ZR];
#pragma unroll
  for (int i = 0; i < ZR; i++) acc[i] = 0.0;
  const bool epi = a.epi_col != nullptr;                     // same fused adjoint epilogue as grad_kernel
  if (epi && threadIdx.x < ZR)
    sRowv[threadIdx.x] = (a.epi_rowv && threadIdx.x < rows) ? a.epi_rowv[(long long)b * a.nA + m0 + threadIdx.x] : 0.0;

  for (int p = 0; p < a.P; p++) {
    const double* h = a.hyp + ((long long)b * a.P + p) * HS;
    const double var = h[0], ls = h[1];
    __syncthreads();
    if (threadIdx.x < ZR) {
      const int i = threadIdx.x;
      const double zt = ((i < rows) ? zrow[m0 + i] : 0.0) / ls;
      sZ[i] = zt; sZ[ZR + i] = __dmul_rn(zt, zt); sZ[2 * ZR + i] = -2.0 * zt;
    }
    const double* fb = nullptr;
    if (mercer) {
      const double* fa = a.featA + ((long long)b * a.P + p) * KP * (long long)a.nA;
      fb = a.featB + ((long long)b * a.P + p) * KP * (long long)a.nB;
      for (int q = threadIdx.x; q < Q; q += ZTHREADS) sW[q] = __dmul_rn(TWO_PI, h[2 + Q + q]);
      for (int idx = threadIdx.x; idx < ZR * 2 * Q; idx += ZTHREADS) {
        const int i = idx / (2 * Q), k = idx - i * 2 * Q;
        sFA[idx] = (i < rows) ? fa[(long long)k * a.nA + m0 + i] : 0.0;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.nB; c += ZTHREADS) {
      const double xt = xrow[c] / ls, xt2 = __dmul_rn(xt, xt);
      const double epi_c = epi ? a.epi_alpha * a.epi_col[(long long)b * a.nB + c] : 0.0;
      const double epi_v = (epi && a.epi_colv) ? a.epi_colv[(long long)b * a.nB + c] : 0.0;
      double kq[ZR], sq[ZR];
#pragma unroll
      for (int i = 0; i < ZR; i++) kq[i] = sq[i] = 0.0;
      if (mercer)
        for (int q = 0; q < Q; q++) {
          const double xc = fb[(long long)q * a.nB + c], xs = fb[(long long)(Q + q) * a.nB + c], w = sW[q];
#pragma unroll
          for (int i = 0; i < ZR; i++) {
            const double zc = sFA[i * 2 * Q + q], zs = sFA[i * 2 * Q + Q + q];
            kq[i] += fma(zc, xc, zs * xs);                     // e_q cos(w_q (z - x))
            sq[i] = fma(w, fma(zs, xc, -zc * xs), sq[i]);      // w_q e_q sin(w_q (z - x))
          }
        }
#pragma unroll
      for (int i = 0; i < ZR; i++) {
        if (i >= rows) break;
        double kb = Kb[(long long)(m0 + i) * a.ldk + c];
        if (epi) kb = fma(epi_c, kb, sRowv[i] * epi_v);
        const double dt = sZ[i] - xt;
        double s;
        if (a.mode == DIST_REFERENCE) s = sqdist_ref(sZ[2 * ZR + i], sZ[ZR + i], xt, xt2);
        else s = dt * dt;
        double rinv;
        const double r = sqrt_pos_rinv(s + 1e-12, rinv);
        if (mercer) acc[i] -= kb * var * exp_neg(r, sT) * fma(kq[i], dt * rinv / ls, sq[i]);
        else acc[i] -= kb * 3.0 * var * exp_neg(1.7320508075688772 * r, sT) * dt / ls;
      }
    }
  }
  block_sum_many<ZR>(acc, ZR, sRed, [&](int j, double tot) {
    if (j < rows) dpts[(long long)b * a.nA + m0 + j] = tot;
  });
}

int launch_kernel_grad_points(const KernArgs& a, double* dpts, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0) return GPX_OK;
  if (a.batch > 65535 || a.P < 1 || !dpts || a.kind == KIND_DIFF_M12 || a.kind == KIND_DIFF_M32) return GPX_ERR_ARG;
  if (a.kind == KIND_MERCER_M12 && (!a.featA || !a.featB || a.Q < 1)) return GPX_ERR_ARG;
  if (a.nB <= 0) return cudaMemsetAsync(dpts, 0, sizeof(double) * (size_t)a.batch * a.nA, st) == cudaSuccess ? GPX_OK : GPX_ERR_LAUNCH;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  const int Qs = a.kind == KIND_MERCER_M12 ? a.Q : 0;
  size_t smem = ((size_t)64 + 3 * ZR + Qs + (size_t)ZR * 2 * Qs) * sizeof(double);
  if (smem > 48 * 1024) return GPX_ERR_ARG;
  dim3 grid((a.nA + ZR - 1) / ZR, a.batch);
  grad_points_kernel<<<grid, ZTHREADS, smem, st>>>(a, dpts);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

}  // namespace gpx

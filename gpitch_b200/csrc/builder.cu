#include "builder.cuh"
#include "fastmath.cuh"
#include <cmath>

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
// Builder: K[b] = sum_p k_p(ptsA, ptsB).  CTA tile 32 x 128, 8 warps (2 x 4), warp tile 16 x 32.
// The rank-2Q feature contraction runs on the FP64 tensor pipe; exp / sqrt / distance on the FP64 ALUs.
// ---------------------------------------------------------------------------------------------------------
constexpr int BBM = 32, BBN = 128, BTHREADS = 256, BMT = 2;
constexpr int B_LDA = BBM + 4, B_LDB = BBN + 4;

__global__ void __launch_bounds__(BTHREADS) build_kernel(const KernArgs a, const int rt_per_cta) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * BBN;
  const int Q = a.Q, HS = 2 + 2 * Q;
  const int KP = (a.kind == KIND_MERCER_M12) ? (2 * Q + 3) / 4 * 4 : 0;
  double* sFA = sm;                    // [KP][B_LDA]
  double* sFB = sFA + KP * B_LDA;      // [KP][B_LDB]
  double* sZ = sFB + KP * B_LDB;       // per-row: raw z, zt (scaled), zt^2, -2 zt     [4][BBM]
  double* sX = sZ + 4 * BBM;           // per-col: raw x, xt, xt^2                     [3][BBN]
  double* sH = sX + 3 * BBN;           // hypers of the current component [HS]
  double* sT = sH + HS;                // 2^(j/64) table for exp_neg
  load_exp_table(sT);

  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * (8 * BMT), wn0 = (warp & 3) * 32;
  double* Kg = a.K + (long long)b * a.sK;
  const bool vec = ((a.ldk & 1) == 0) && ((((uintptr_t)Kg) & 15) == 0);

  // One CTA owns a 128-column strip and walks rt_per_cta row tiles of 32 inducing points: the column-side data
  // (scaled inputs and, for a single-component kernel, the Mercer feature tile) is staged once per strip.
  const int n_rt = (a.nA + BBM - 1) / BBM;
  const int rt0 = blockIdx.y * rt_per_cta, rt1 = min(n_rt, rt0 + rt_per_cta);
  bool x_staged = false;

  for (int rt = rt0; rt < rt1; rt++) {
    const int m0 = rt * BBM;
    double tot[BMT][4][2];
#pragma unroll
    for (int i = 0; i < BMT; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) tot[i][j][0] = tot[i][j][1] = 0.0;

    for (int p = 0; p < a.P; p++) {
      __syncthreads();  // previous component / row tile fully consumed
      const double* h = a.hyp + ((long long)b * a.P + p) * HS;
      for (int i = threadIdx.x; i < HS; i += BTHREADS) sH[i] = h[i];
      const double ls = h[1];
      for (int i = threadIdx.x; i < BBM; i += BTHREADS) {
        const int r = m0 + i;
        const double z = (r < a.nA) ? zrow[r] : 0.0;
        const double zt = z / ls;
        sZ[i] = z; sZ[BBM + i] = zt; sZ[2 * BBM + i] = __dmul_rn(zt, zt); sZ[3 * BBM + i] = -2.0 * zt;
      }
      const bool stage_x = !(a.P == 1 && x_staged);
      if (stage_x)
        for (int i = threadIdx.x; i < BBN; i += BTHREADS) {
          const int c = n0 + i;
          const double x = (c < a.nB) ? xrow[c] : 0.0;
          const double xt = x / ls;
          sX[i] = x; sX[BBN + i] = xt; sX[2 * BBN + i] = __dmul_rn(xt, xt);
        }
      if (a.kind == KIND_MERCER_M12) {
        const double* fa = a.featA + ((long long)b * a.P + p) * KP * (long long)a.nA;
        for (int idx = threadIdx.x; idx < KP * BBM; idx += BTHREADS) {
          int k = idx / BBM, i = idx - k * BBM;
          sFA[k * B_LDA + i] = (m0 + i < a.nA) ? fa[(long long)k * a.nA + m0 + i] : 0.0;
        }
        if (stage_x) {
          const double* fb = a.featB + ((long long)b * a.P + p) * KP * (long long)a.nB;
          for (int idx = threadIdx.x; idx < KP * BBN; idx += BTHREADS) {
            int k = idx / BBN, i = idx - k * BBN;
            sFB[k * B_LDB + i] = (n0 + i < a.nB) ? fb[(long long)k * a.nB + n0 + i] : 0.0;
          }
        }
      }
      x_staged = true;
      __syncthreads();
      const double var = sH[0];

      double acc[BMT][4][2];
      if (a.kind == KIND_MERCER_M12) {
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
#pragma unroll
      for (int i = 0; i < BMT; i++) {
        const int rl = wm0 + i * 8 + g;
        const double z = sZ[rl], zt = sZ[BBM + rl], zt2 = sZ[2 * BBM + rl], m2zt = sZ[3 * BBM + rl];
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cl = wn0 + j * 8 + 2 * t + e;
            double kv;
            if (a.kind == KIND_DIFF_M12) {
              // Matern12sm.K: r = |z - x + 1e-12|; var * exp(-r/l) * sum_q e_q cos(2 pi f_q r)   (:47-56)
              const double r = fabs(__dadd_rn(__dadd_rn(z, -sX[cl]), 1e-12));
              double k = 0.0;
              for (int q = 0; q < Q; q++) {
                const double ph = __dmul_rn(__dmul_rn(TWO_PI, sH[2 + Q + q]), r);
                const double term = sH[2 + q] * cos(ph);
                k = (q == 0) ? term : k + term;
              }
              kv = (var * exp(-(r / sH[1]))) * k;
            } else {
              const double xt = sX[BBN + cl];
              double s;
              if (a.mode == DIST_REFERENCE) s = sqdist_ref(m2zt, zt2, xt, sX[2 * BBN + cl]);
              else { const double d = zt - xt; s = d * d; }
              const double r = sqrt_pos(s + 1e-12);
              if (a.kind == KIND_MERCER_M12) kv = (var * exp_neg(r, sT)) * acc[i][j][e];
              else { const double s3r = 1.7320508075688772 * r; kv = (var * (1.0 + s3r)) * exp_neg(s3r, sT); }
            }
            tot[i][j][e] = (p == 0) ? kv : tot[i][j][e] + kv;
          }
        }
      }
    }

#pragma unroll
    for (int i = 0; i < BMT; i++) {
      const int row = m0 + wm0 + i * 8 + g;
      if (row >= a.nA) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int col = n0 + wn0 + j * 8 + 2 * t;
        double v0 = tot[i][j][0], v1 = tot[i][j][1];
        if (a.jitter != 0.0) { if (row == col) v0 += a.jitter; if (row == col + 1) v1 += a.jitter; }
        double* dst = Kg + (long long)row * a.ldk + col;
        if (vec && col + 1 < a.nB) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        else { if (col < a.nB) dst[0] = v0; if (col + 1 < a.nB) dst[1] = v1; }
      }
    }
  }
}

int launch_kernel_build(const KernArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0 || a.nB <= 0) return GPX_OK;
  if (a.batch > 65535 || a.P < 1) return GPX_ERR_ARG;
  if (a.kind == KIND_MERCER_M12 && (!a.featA || !a.featB || a.Q < 1)) return GPX_ERR_ARG;
  const int KP = (a.kind == KIND_MERCER_M12) ? feat_rows(a.Q) : 0;
  size_t smem = ((size_t)KP * (B_LDA + B_LDB) + 4 * BBM + 3 * BBN + 2 + 2 * a.Q + 64) * sizeof(double);
  if (smem > 200 * 1024) return GPX_ERR_ARG;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // row tiles per CTA: as many as possible while keeping >= ~6 CTAs per SM in flight
  const int n_rt = (a.nA + BBM - 1) / BBM, strips = (a.nB + BBN - 1) / BBN;
  long long tiles = (long long)n_rt * strips * a.batch;
  int rt_per = (int)(tiles / 900);
  rt_per = rt_per < 1 ? 1 : (rt_per > n_rt ? n_rt : rt_per);
  dim3 grid(strips, (n_rt + rt_per - 1) / rt_per, a.batch);
  build_kernel<<<grid, BTHREADS, smem, st>>>(a, rt_per);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Hyper-parameter gradient: dhyp[b,p,:] += sum_{m,n} Kbar[m,n] dK_p[m,n]/dtheta   (SURVEY Appendix B.1).
// K is never re-read: every term is re-evaluated from the points / features, Kbar is read exactly once.
// One thread per column, GBM rows per CTA; rows are processed four at a time with the four Kbar loads issued
// up front (the loop is otherwise one global-load latency per row); per-row quantities (z/l, (z/l)^2, features)
// are staged in shared memory; register accumulators -> block reduce -> one atomic per value.
// ---------------------------------------------------------------------------------------------------------
constexpr int GBM = 40, GTHREADS = 256, GROWS = 4;

template <int GQ>
struct GradAcc {
  double e[GQ > 0 ? GQ : 1], f[GQ > 0 ? GQ : 1];
};

template <bool NEED_EF, int GQ>
__global__ void __launch_bounds__(GTHREADS, 2) grad_kernel(const KernArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[32];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * GBM, c = blockIdx.x * GTHREADS + threadIdx.x;
  const int Q = a.Q, HS = 2 + 2 * Q;
  const int KP = (a.kind == KIND_MERCER_M12) ? (2 * Q + 3) / 4 * 4 : 0;
  double* sT = sm;                  // exp table [64]
  double* sZ = sT + 64;             // per row: z, z/l, (z/l)^2, -2 z/l   [4][GBM]
  double* sFA = sZ + 4 * GBM;       // [GBM][2Q] row features (row-major per inducing point -> broadcast reads)
  load_exp_table(sT);
  const double* zrow = a.ptsA + (long long)(b / a.divA) * a.nA;
  const double* xrow = a.ptsB + (long long)(b / a.divB) * a.nB;
  const double* Kb = a.K + (long long)b * a.sK;
  const bool colv = c < a.nB;
  const int cc = colv ? c : 0;
  const double x = colv ? xrow[c] : 0.0;
  const int rows = min(GBM, a.nA - m0);

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
      for (int idx = threadIdx.x; idx < GBM * 2 * Q; idx += GTHREADS) {
        int k = idx / GBM, i = idx - k * GBM;
        sFA[i * 2 * Q + k] = (i < rows) ? fa[(long long)k * a.nA + m0 + i] : 0.0;
      }
    }
    __syncthreads();
    const double xt = x / ls, xt2 = __dmul_rn(xt, xt);
    double a_var = 0.0, a_len = 0.0;

    if (a.kind == KIND_DIFF_M12) {
      // r = |z - x + 1e-12|, K = var exp(-r/l) sum_q e_q cos(w_q r);  dK/dl = K r / l^2   (not on the named path:
      // plain libm trig per element)
      for (int q0 = 0; q0 < Q || q0 == 0; q0 += (GQ > 0 ? GQ : 1)) {
        GradAcc<GQ> A;
#pragma unroll
        for (int q = 0; q < (GQ > 0 ? GQ : 1); q++) A.e[q] = A.f[q] = 0.0;
        if (colv)
          for (int i = 0; i < rows; i++) {
            const double r = fabs(__dadd_rn(__dadd_rn(sZ[i], -x), 1e-12));
            const double W = Kb[(long long)(m0 + i) * a.ldk + c] * exp(-(r / ls));
            double k = 0.0;
            for (int q = 0; q < Q; q++) {
              double sn, cs;
              sincos(__dmul_rn(__dmul_rn(TWO_PI, h[2 + Q + q]), r), &sn, &cs);
              k += h[2 + q] * cs;
              if (NEED_EF && q >= q0 && q < q0 + GQ) { A.e[q - q0] += W * cs; A.f[q - q0] += W * r * sn; }
            }
            if (q0 == 0) { a_var += W * k; a_len += W * k * r; }
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
    const int nchunk = (mercer && NEED_EF && GQ > 0) ? (Q + GQ1 - 1) / GQ1 : 1;
    for (int ch = 0; ch < nchunk; ch++) {
      const int q0 = ch * (GQ > 0 ? GQ : 1);
      double xc[GQ > 0 ? GQ : 1], xs[GQ > 0 ? GQ : 1];
      GradAcc<GQ> A;
#pragma unroll
      for (int q = 0; q < (GQ > 0 ? GQ : 1); q++) {
        const bool v = mercer && NEED_EF && (q0 + q < Q);
        xc[q] = v ? fb[(long long)(q0 + q) * a.nB + cc] : 0.0;
        xs[q] = v ? fb[(long long)(Q + q0 + q) * a.nB + cc] : 0.0;
        A.e[q] = A.f[q] = 0.0;
      }
      for (int i0 = 0; i0 < rows; i0 += GROWS) {
        double kb[GROWS];
#pragma unroll
        for (int u = 0; u < GROWS; u++)
          kb[u] = (colv && i0 + u < rows) ? __ldg(Kb + (long long)(m0 + i0 + u) * a.ldk + c) : 0.0;
#pragma unroll
        for (int u = 0; u < GROWS; u++) {
          const int i = i0 + u;                                  // rows beyond `rows` carry kb = 0 -> no effect
          const double z = sZ[i], zt = sZ[GBM + i];
          double s;
          if (a.mode == DIST_REFERENCE) s = sqdist_ref(sZ[3 * GBM + i], sZ[2 * GBM + i], xt, xt2);
          else { const double d = zt - xt; s = d * d; }
          const double r = sqrt_pos(s + 1e-12);
          if (!mercer) {                                         // Matern-3/2: dK/dvar = K/var, dK/dl = 3 var E s / l
            const double s3r = 1.7320508075688772 * r, E = exp_neg(s3r, sT);
            a_var += kb[u] * (1.0 + s3r) * E;
            a_len += kb[u] * E * s;
            continue;
          }
          const double W = kb[u] * exp_neg(r, sT);
          const double* fz = sFA + i * 2 * Q;
          if (NEED_EF && GQ > 0) {
            const double Wd = W * (z - x);
            double k = 0.0;
#pragma unroll
            for (int q = 0; q < GQ; q++) {
              if (q0 + q < Q) {
                const double zc = fz[q0 + q], zs = fz[Q + q0 + q];
                const double cq = fma(zc, xc[q], zs * xs[q]);     // e_q cos(w_q (z - x))
                const double sq = fma(zs, xc[q], -zc * xs[q]);    // e_q sin(w_q (z - x))
                k += cq;
                A.e[q] = fma(W, cq, A.e[q]);
                A.f[q] = fma(Wd, sq, A.f[q]);
              }
            }
            const double Wk = W * k;
            a_var += Wk;
            a_len = fma(Wk, s / r, a_len);
          } else {                                               // energies / frequencies fixed: only k is needed
            double k = 0.0;
            for (int q = 0; q < Q; q++)
              k += fz[q] * fb[(long long)q * a.nB + cc] + fz[Q + q] * fb[(long long)(Q + q) * a.nB + cc];
            a_var += W * k;
            a_len += W * k * (s / r);
          }
        }
      }
      if (mercer && NEED_EF && GQ > 0)
        for (int q = 0; q < GQ && q0 + q < Q; q++) {
          double se = block_sum<false>(A.e[q], red), sf = block_sum<false>(A.f[q], red);
          if (threadIdx.x == 0) {
            const double eq = h[2 + q0 + q];
            atomicAdd(dh + 2 + q0 + q, eq > 0.0 ? var * se / eq : 0.0);        // dK/de_q = var E cos
            atomicAdd(dh + 2 + Q + q0 + q, -var * TWO_PI * sf);                 // dK/df_q = -var E e_q 2 pi d sin
          }
        }
      if (ch > 0) { a_var = 0.0; a_len = 0.0; }      // var / len sums are complete after the first chunk
      if (ch == 0) {
        const double sv = block_sum<false>(a_var, red), sl = block_sum<false>(a_len, red);
        if (threadIdx.x == 0) {
          if (mercer) { atomicAdd(dh + 0, sv); atomicAdd(dh + 1, var * sl / ls); }
          else { atomicAdd(dh + 0, sv); atomicAdd(dh + 1, 3.0 * var * sl / ls); }
        }
        a_var = 0.0; a_len = 0.0;
      }
    }
  }
}

template <bool NEED_EF, int GQ>
static int launch_grad_cfg(const KernArgs& a, cudaStream_t st) {
  size_t smem = ((size_t)64 + 4 * GBM + (size_t)GBM * 2 * a.Q) * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(grad_kernel<NEED_EF, GQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.nB + GTHREADS - 1) / GTHREADS, (a.nA + GBM - 1) / GBM, a.batch);
  grad_kernel<NEED_EF, GQ><<<grid, GTHREADS, smem, st>>>(a);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

int launch_kernel_grad(const KernArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.nA <= 0 || a.nB <= 0) return GPX_OK;
  if (a.batch > 65535 || a.P < 1 || !a.dhyp) return GPX_ERR_ARG;
  if (a.kind == KIND_MERCER_M12 && (!a.featA || !a.featB || a.Q < 1)) return GPX_ERR_ARG;
  if (init_fastmath() != GPX_OK) return GPX_ERR_LAUNCH;
  if (!a.need_ef || a.kind == KIND_MATERN32) return launch_grad_cfg<false, 0>(a, st);
  if (a.Q <= 4) return launch_grad_cfg<true, 4>(a, st);
  if (a.Q <= 6) return launch_grad_cfg<true, 6>(a, st);
  return launch_grad_cfg<true, 10>(a, st);          // Q <= 10 in one pass; larger Q in chunks of 10 partials
}

}  // namespace gpx

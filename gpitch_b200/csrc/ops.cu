#include "ops.cuh"

namespace gpx {

// ------------------------------------------------------------------------------------------ column statistics
template <int MODE, bool HAS_B>
__global__ void __launch_bounds__(256) cond_colstats_kernel(const double* __restrict__ A, const double* __restrict__ LTA,
                                                            long long sA, int ld, const double* __restrict__ mu,
                                                            const double* __restrict__ kdiag, double* __restrict__ fmean,
                                                            double* __restrict__ fvar, int M, int N) {
  extern __shared__ double smu[];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < M; i += blockDim.x) smu[i] = mu[(long long)b * M + i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double* Ab = A + (long long)b * sA + n;
  const double* Lb = HAS_B ? LTA + (long long)b * sA + n : nullptr;
  double m0 = 0, m1 = 0, a0 = 0, a1 = 0, l0 = 0, l1 = 0;
  int m = 0;
  for (; m + 1 < M; m += 2) {
    const double x0 = Ab[(long long)m * ld], x1 = Ab[(long long)(m + 1) * ld];
    m0 += x0 * smu[m]; m1 += x1 * smu[m + 1];
    a0 += x0 * x0; a1 += x1 * x1;
    if (HAS_B) {
      const double y0 = Lb[(long long)m * ld], y1 = Lb[(long long)(m + 1) * ld];
      if (MODE == 0) { l0 += y0 * y0; l1 += y1 * y1; } else { l0 += x0 * y0; l1 += x1 * y1; }
    }
  }
  if (m < M) {
    const double x0 = Ab[(long long)m * ld];
    m0 += x0 * smu[m]; a0 += x0 * x0;
    if (HAS_B) { const double y0 = Lb[(long long)m * ld]; l0 += (MODE == 0) ? y0 * y0 : x0 * y0; }
  }
  fmean[(long long)b * N + n] = m0 + m1;
  // mode 0: Kdiag - sum A^2 + sum LTA^2 (GPflow conditional);  mode 1: Kdiag + sum A o B (G-form, A = Kmn, B = G Kmn)
  fvar[(long long)b * N + n] = (MODE == 0) ? (kdiag[b] - (a0 + a1)) + (l0 + l1) : kdiag[b] + (l0 + l1);
}

// Same statistics for launches too small to fill the GPU (one window: N / 256 CTAs of serial M-long load chains, 25 us for a
// 200 x 200 matrix): 32 columns x 8 row groups per CTA, four rows in flight per thread, fixed-order reduction over the row
// groups in shared memory -- 8 x the CTAs and a 32 x shorter dependent chain.
template <int MODE, bool HAS_B>
__global__ void __launch_bounds__(256) cond_colstats_small_kernel(const double* __restrict__ A, const double* __restrict__ LTA,
                                                                  long long sA, int ld, const double* __restrict__ mu,
                                                                  const double* __restrict__ kdiag, double* __restrict__ fmean,
                                                                  double* __restrict__ fvar, int M, int N) {
  extern __shared__ double smu[];
  double* red = smu + M;                               // [3][8][32]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < M; i += blockDim.x) smu[i] = mu[(long long)b * M + i];
  __syncthreads();
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  double mm = 0.0, aa = 0.0, ll = 0.0;
  if (n < N) {
    const double* Ab = A + (long long)b * sA + n;
    const double* Lb = HAS_B ? LTA + (long long)b * sA + n : nullptr;
    for (int m0 = rg; m0 < M; m0 += 32) {
      double x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int m = m0 + 8 * u;
        x[u] = (m < M) ? Ab[(long long)m * ld] : 0.0;
        y[u] = (HAS_B && m < M) ? Lb[(long long)m * ld] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int m = m0 + 8 * u;
        if (m < M) {
          mm += x[u] * smu[m];
          aa += x[u] * x[u];
          if (HAS_B) ll += (MODE == 0) ? y[u] * y[u] : x[u] * y[u];
        }
      }
    }
  }
  red[(0 * 8 + rg) * 32 + c] = mm;
  red[(1 * 8 + rg) * 32 + c] = aa;
  red[(2 * 8 + rg) * 32 + c] = ll;
  __syncthreads();
  if (rg == 0 && n < N) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) { s0 += red[(0 * 8 + r) * 32 + c]; s1 += red[(1 * 8 + r) * 32 + c]; s2 += red[(2 * 8 + r) * 32 + c]; }
    fmean[(long long)b * N + n] = s0;
    fvar[(long long)b * N + n] = (MODE == 0) ? (kdiag[b] - s1) + s2 : kdiag[b] + s2;
  }
}

int launch_cond_colstats(const double* A, const double* LTA, long long sA, int ld, const double* mu,
                         const double* kdiag, double* fmean, double* fvar, int M, int N, int batch, int mode,
                         cudaStream_t st) {
  if (batch <= 0 || N <= 0) return GPX_OK;
  if (batch > 65535 || M * sizeof(double) > 40 * 1024) return GPX_ERR_ARG;
  if ((long long)((N + 255) / 256) * batch < 148) {      // latency-bound launch: the many-CTA variant
    dim3 grid((N + 31) / 32, batch);
    const size_t sm = (M + 3 * 8 * 32) * sizeof(double);
    if (mode == 1) cond_colstats_small_kernel<1, true><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
    else if (LTA) cond_colstats_small_kernel<0, true><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
    else cond_colstats_small_kernel<0, false><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
    GPX_CHECK_LAUNCH();
    return GPX_OK;
  }
  dim3 grid((N + 255) / 256, batch);
  const size_t sm = M * sizeof(double);
  if (mode == 1) cond_colstats_kernel<1, true><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
  else if (LTA) cond_colstats_kernel<0, true><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
  else cond_colstats_kernel<0, false><<<grid, 256, sm, st>>>(A, LTA, sA, ld, mu, kdiag, fmean, fvar, M, N);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ scaled rank-1 update
// out[b,m,n] = alpha * colscale[b,n] * T[b,m,n] + rowvec[b,m] * colvec[b,n]   (Kbar_mn = 2 T diag(vbar) + a mbar^T)
__global__ void __launch_bounds__(256) scale_rank1_kernel(const double* __restrict__ T, long long sT, int ld,
                                                          const double* __restrict__ cs, const double* __restrict__ rv,
                                                          const double* __restrict__ cv, double alpha,
                                                          double* __restrict__ out, int M, int N) {
  const int b = blockIdx.z;
  const int n2 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (n2 >= N) return;
  const bool two = n2 + 1 < N;
  const double c0 = alpha * cs[(long long)b * N + n2], c1 = two ? alpha * cs[(long long)b * N + n2 + 1] : 0.0;
  const double v0 = cv[(long long)b * N + n2], v1 = two ? cv[(long long)b * N + n2 + 1] : 0.0;
  const bool vec = two && ((ld & 1) == 0);
  for (int m = blockIdx.y; m < M; m += gridDim.y) {
    const double r = rv[(long long)b * M + m];
    const long long off = (long long)b * sT + (long long)m * ld + n2;
    if (vec) {
      const double2 t = *reinterpret_cast<const double2*>(T + off);
      *reinterpret_cast<double2*>(out + off) = make_double2(fma(c0, t.x, r * v0), fma(c1, t.y, r * v1));
    } else {
      out[off] = fma(c0, T[off], r * v0);
      if (two) out[off + 1] = fma(c1, T[off + 1], r * v1);
    }
  }
}

int launch_scale_rank1(const double* T, long long sT, int ld, const double* cs, const double* rv, const double* cv,
                       double alpha, double* out, int M, int N, int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0 || N <= 0) return GPX_OK;
  if (batch > 65535) return GPX_ERR_ARG;
  dim3 grid(((N + 1) / 2 + 255) / 256, M < 50 ? M : 50, batch);
  scale_rank1_kernel<<<grid, 256, 0, st>>>(T, sT, ld, cs, rv, cv, alpha, out, M, N);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ row dots
__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ A, long long sA, int ld,
                                                     const double* __restrict__ v, long long sV, double* __restrict__ out,
                                                     int M, int N) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const double* row = A + (long long)b * sA + (long long)m * ld;
  const double* vb = v + (long long)b * sV;
  double s0 = 0, s1 = 0;
  int n = lane;
  for (; n + 224 < N; n += 256) {          // eight loads in flight per lane (single-window launches are latency-bound)
    double r[8], w[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { r[u] = row[n + 32 * u]; w[u] = vb[n + 32 * u]; }
#pragma unroll
    for (int u = 0; u < 8; u += 2) { s0 += r[u] * w[u]; s1 += r[u + 1] * w[u + 1]; }
  }
  for (; n + 32 < N; n += 64) { s0 += row[n] * vb[n]; s1 += row[n + 32] * vb[n + 32]; }
  if (n < N) s0 += row[n] * vb[n];
  const double s = warp_sum(s0 + s1);
  if (lane == 0) out[(long long)b * M + m] = s;
}

int launch_rowdot(const double* A, long long sA, int ld, const double* v, long long sV, double* out, int M, int N,
                  int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  if (batch > 65535) return GPX_ERR_ARG;
  dim3 grid((M + 7) / 8, batch);
  rowdot_kernel<<<grid, 256, 0, st>>>(A, sA, ld, v, sV, out, M, N);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ Gauss-Hermite
// np.polynomial.hermite.hermgauss(20) (GPflow quadrature.hermgauss), weights already divided by sqrt(pi)
// (likelihoods.py:35-37).  Filled at load time by gpx_set_hermgauss() from the host's NumPy so that the nodes are
// bit-identical to the reference's.
__constant__ double c_ghx[64];
__constant__ double c_ghw[64];
__constant__ int c_ghn = 0;

int set_hermgauss(const double* x, const double* w, int n) {
  if (n < 1 || n > 64) return GPX_ERR_ARG;
  if (cudaMemcpyToSymbol(c_ghx, x, n * sizeof(double)) != cudaSuccess) return GPX_ERR_LAUNCH;
  if (cudaMemcpyToSymbol(c_ghw, w, n * sizeof(double)) != cudaSuccess) return GPX_ERR_LAUNCH;
  if (cudaMemcpyToSymbol(c_ghn, &n, sizeof(int)) != cudaSuccess) return GPX_ERR_LAUNCH;
  return GPX_OK;
}

#define PI_D 3.141592653589793

template <int NLIN>
__device__ __forceinline__ void nlin_eval(double X, double& s, double& ds) {
  if (NLIN == 0) {         // logistic_tf: 1 / (1 + exp(-2 (x - pi)))            (methods.py:216-218)
    s = 1.0 / (1.0 + exp(-2.0 * (X - PI_D)));
    ds = 2.0 * s * (1.0 - s);
  } else if (NLIN == 1) {  // softplus_tf: log(exp(x) + 1)                       (methods.py:220-222)
    const double ex = exp(X);
    s = log(ex + 1.0);
    ds = 1.0 / (1.0 + exp(-X));
  } else {                 // gaussfun_tf: exp(-2 (x - pi)^2)                    (methods.py:232-233)
    const double d = X - PI_D;
    s = exp(-2.0 * (d * d));
    ds = s * (-4.0 * d);
  }
}

template <int NLIN>
__global__ void __launch_bounds__(256) varexp_kernel(const double* __restrict__ Fmu, const double* __restrict__ Fvar,
                                                     const double* __restrict__ Y, const double* __restrict__ noise,
                                                     int P, int W, int N, double* __restrict__ ve_sum,
                                                     double* __restrict__ dFmu, double* __restrict__ dFvar,
                                                     double* __restrict__ dnoise, double* __restrict__ ve_pt) {
  __shared__ double red[32];
  const int w = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < N;
  const double s2 = noise[w];
  const long long WN = (long long)N;                      // stride between the 2P latent rows of one window
  const long long fbase = (long long)w * 2 * P * N + n;   // Fmu / Fvar [W, 2P, N]
  const long long base = (long long)w * N + n;            // Y / ve_pointwise [W, N]
  const int H = c_ghn;
  double ve = 0.0, dn = 0.0;
  if (valid) {
    const double y = Y[base];
    // pass 1: S = sum_i a_i, Bt = sum_i E2_i (vf_i + mf_i^2), sum a_i^2
    double S = 0.0, Bt = 0.0, Saa = 0.0;
    for (int i = 0; i < P; i++) {
      const double mg = Fmu[i * WN + fbase], vg = Fvar[i * WN + fbase];
      const double mf = Fmu[(P + i) * WN + fbase], vf = Fvar[(P + i) * WN + fbase];
      const double sd = sqrt(2.0 * vg);
      double E1 = 0.0, E2 = 0.0;
      for (int h = 0; h < H; h++) {
        double s, ds;
        nlin_eval<NLIN>(c_ghx[h] * sd + mg, s, ds);
        E1 += c_ghw[h] * s;
        E2 += c_ghw[h] * (s * s);
      }
      const double a = E1 * mf;
      S += a; Saa += a * a; Bt += E2 * (vf + mf * mf);
    }
    const double C = (P == 1) ? 0.0 : (S * S - Saa);
    const double quad = (y * y - 2.0 * y * S + Bt) + C;
    ve = -0.5 * ((1.0 / s2) * quad + 1.8378770664093453 + log(s2));   // log(2 pi)
    dn = 0.5 * quad / (s2 * s2) - 0.5 / s2;
    if (ve_pt) ve_pt[base] = ve;
    // pass 2: gradients (quadrature re-evaluated; 20 P exps per sample either way, no [n,20] temporaries)
    if (dFmu) {
      for (int i = 0; i < P; i++) {
        const double mg = Fmu[i * WN + fbase], vg = Fvar[i * WN + fbase];
        const double mf = Fmu[(P + i) * WN + fbase], vf = Fvar[(P + i) * WN + fbase];
        const double sd = sqrt(2.0 * vg);
        double E1 = 0.0, E2 = 0.0, d1m = 0.0, d1v = 0.0, d2m = 0.0, d2v = 0.0;
        for (int h = 0; h < H; h++) {
          double s, ds;
          const double xh = c_ghx[h], wh = c_ghw[h];
          nlin_eval<NLIN>(xh * sd + mg, s, ds);
          E1 += wh * s; E2 += wh * (s * s);
          const double wds = wh * ds, w2 = 2.0 * s * wds;
          d1m += wds; d1v += wds * xh; d2m += w2; d2v += w2 * xh;
        }
        const double inv_sd = (sd > 0.0) ? 1.0 / sd : 0.0;   // d sqrt(2 vg)/d vg = 1/sd
        d1v *= inv_sd; d2v *= inv_sd;
        const double a = E1 * mf;
        const double da = (y - S + a) / s2;              // d var_exp / d a_i
        const double dE1 = da * mf;
        const double dE2 = -0.5 * (vf + mf * mf) / s2;
        dFmu[i * WN + fbase] = dE1 * d1m + dE2 * d2m;
        dFvar[i * WN + fbase] = dE1 * d1v + dE2 * d2v;
        dFmu[(P + i) * WN + fbase] = da * E1 - E2 * mf / s2;
        dFvar[(P + i) * WN + fbase] = -0.5 * E2 / s2;
      }
    }
  }
  ve = block_sum<false>(ve, red);
  dn = block_sum<false>(dn, red);
  if (threadIdx.x == 0) {
    atomicAdd(ve_sum + w, ve);
    if (dnoise) atomicAdd(dnoise + w, dn);
  }
}

// Single-pass variant for P <= VE_PMAX sources (every Pdgp configuration of the demos and of BASELINE): the quadrature of
// source i is evaluated ONCE; the outputs that depend on S = sum_i a_i only through da_i = (y - S + a_i) / s2 are written
// provisionally with S = 0 and their three S-coefficients (mf d1m, mf d1v, E1) / s2 stay in shared memory until S is
// known -- 20 P exponentials per sample instead of 40 P.
constexpr int VE_PMAX = 16, VE_THREADS = 128;

template <int NLIN>
__global__ void __launch_bounds__(VE_THREADS) varexp_onepass_kernel(const double* __restrict__ Fmu,
                                                                    const double* __restrict__ Fvar,
                                                                    const double* __restrict__ Y,
                                                                    const double* __restrict__ noise, int P, int W, int N,
                                                                    double* __restrict__ ve_sum, double* __restrict__ dFmu,
                                                                    double* __restrict__ dFvar, double* __restrict__ dnoise,
                                                                    double* __restrict__ ve_pt) {
  extern __shared__ __align__(16) double sco[];          // [P][3][VE_THREADS]
  __shared__ double red[32];
  const int w = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < N;
  const double s2 = noise[w], is2 = 1.0 / s2;
  const long long WN = (long long)N;
  const long long fbase = (long long)w * 2 * P * N + n;
  const long long base = (long long)w * N + n;
  const int H = c_ghn;
  double ve = 0.0, dn = 0.0;
  if (valid) {
    const double y = Y[base];
    double S = 0.0, Bt = 0.0, Saa = 0.0;
    for (int i = 0; i < P; i++) {
      const double mg = Fmu[i * WN + fbase], vg = Fvar[i * WN + fbase];
      const double mf = Fmu[(P + i) * WN + fbase], vf = Fvar[(P + i) * WN + fbase];
      const double sd = sqrt(2.0 * vg);
      double E1 = 0.0, E2 = 0.0, d1m = 0.0, d1v = 0.0, d2m = 0.0, d2v = 0.0;
      for (int h = 0; h < H; h++) {
        double sg, ds;
        const double xh = c_ghx[h], wh = c_ghw[h];
        nlin_eval<NLIN>(xh * sd + mg, sg, ds);
        E1 += wh * sg; E2 += wh * (sg * sg);
        const double wds = wh * ds, w2 = 2.0 * sg * wds;
        d1m += wds; d1v += wds * xh; d2m += w2; d2v += w2 * xh;
      }
      const double inv_sd = (sd > 0.0) ? 1.0 / sd : 0.0;
      d1v *= inv_sd; d2v *= inv_sd;
      const double a = E1 * mf;
      S += a; Saa += a * a; Bt += E2 * (vf + mf * mf);
      if (dFmu) {
        const double da0 = (y + a) * is2;                  // da with S = 0
        const double dE2 = -0.5 * (vf + mf * mf) * is2;
        dFmu[i * WN + fbase] = (da0 * mf) * d1m + dE2 * d2m;
        dFvar[i * WN + fbase] = (da0 * mf) * d1v + dE2 * d2v;
        dFmu[(P + i) * WN + fbase] = da0 * E1 - E2 * mf * is2;
        dFvar[(P + i) * WN + fbase] = -0.5 * E2 * is2;
        sco[(i * 3 + 0) * VE_THREADS + threadIdx.x] = mf * d1m * is2;
        sco[(i * 3 + 1) * VE_THREADS + threadIdx.x] = mf * d1v * is2;
        sco[(i * 3 + 2) * VE_THREADS + threadIdx.x] = E1 * is2;
      }
    }
    const double C = (P == 1) ? 0.0 : (S * S - Saa);
    const double quad = (y * y - 2.0 * y * S + Bt) + C;
    ve = -0.5 * (is2 * quad + 1.8378770664093453 + log(s2));
    dn = 0.5 * quad / (s2 * s2) - 0.5 / s2;
    if (ve_pt) ve_pt[base] = ve;
    if (dFmu) {                                            // fix-up: the -S / s2 part of da_i (own writes: no barrier needed)
      for (int i = 0; i < P; i++) {
        dFmu[i * WN + fbase] -= S * sco[(i * 3 + 0) * VE_THREADS + threadIdx.x];
        dFvar[i * WN + fbase] -= S * sco[(i * 3 + 1) * VE_THREADS + threadIdx.x];
        dFmu[(P + i) * WN + fbase] -= S * sco[(i * 3 + 2) * VE_THREADS + threadIdx.x];
      }
    }
  }
  ve = block_sum<false>(ve, red);
  dn = block_sum<false>(dn, red);
  if (threadIdx.x == 0) {
    atomicAdd(ve_sum + w, ve);
    if (dnoise) atomicAdd(dnoise + w, dn);
  }
}

int launch_varexp(const double* Fmu, const double* Fvar, const double* Y, const double* noise, int P, int W, int N,
                  int nlin, double* ve_sum, double* dFmu, double* dFvar, double* dnoise, double* ve_pt,
                  cudaStream_t st) {
  if (W <= 0 || N <= 0) return GPX_OK;
  if (W > 65535 || P < 1 || nlin < 0 || nlin > 2) return GPX_ERR_ARG;
  if (dFmu && P <= VE_PMAX) {          // gradients wanted, few sources: one quadrature pass
    dim3 g1((N + VE_THREADS - 1) / VE_THREADS, W);
    const size_t smem = (size_t)P * 3 * VE_THREADS * sizeof(double);
    if (nlin == 0) varexp_onepass_kernel<0><<<g1, VE_THREADS, smem, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
    else if (nlin == 1) varexp_onepass_kernel<1><<<g1, VE_THREADS, smem, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
    else varexp_onepass_kernel<2><<<g1, VE_THREADS, smem, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
    GPX_CHECK_LAUNCH();
    return GPX_OK;
  }
  dim3 grid((N + 255) / 256, W);
  if (nlin == 0) varexp_kernel<0><<<grid, 256, 0, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
  else if (nlin == 1) varexp_kernel<1><<<grid, 256, 0, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
  else varexp_kernel<2><<<grid, 256, 0, st>>>(Fmu, Fvar, Y, noise, P, W, N, ve_sum, dFmu, dFvar, dnoise, ve_pt);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ whitened KL
// One CTA of 1024 threads per latent GP, one warp per row (no index division, upper triangle never read).
__global__ void __launch_bounds__(1024) gauss_kl_white_kernel(const double* __restrict__ q_mu,
                                                              const double* __restrict__ q_sqrt, int M,
                                                              double* __restrict__ kl, double* __restrict__ dmu,
                                                              double* __restrict__ dLq, double* __restrict__ tril_out) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  const double* mu = q_mu + (long long)b * M;
  const double* Lq = q_sqrt + (long long)b * M * M;
  double* dL = dLq ? dLq + (long long)b * M * M : nullptr;
  double* tl = tril_out ? tril_out + (long long)b * M * M : nullptr;     // tf.matrix_band_part(q_sqrt, -1, 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double m = mu[i], d = Lq[(long long)i * M + i];
    acc += m * m - log(d * d);
    if (dmu) dmu[(long long)b * M + i] = m;
  }
  for (int i = warp; i < M; i += nw) {
    const double* row = Lq + (long long)i * M;
    for (int j = lane; j < M; j += 32) {
      double g = 0.0, v = 0.0;
      if (j <= i) {
        v = row[j];
        acc += v * v;
        g = (i == j) ? v - 1.0 / v : v;
      }
      if (dL) dL[(long long)i * M + j] = g;
      if (tl) tl[(long long)i * M + j] = v;
    }
  }
  acc = block_sum<false>(acc, red);
  if (threadIdx.x == 0) kl[b] = 0.5 * (acc - (double)M);
}

int launch_gauss_kl_white(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* dmu,
                          double* dLq, cudaStream_t st, double* tril_out) {
  if (batch <= 0) return GPX_OK;
  gauss_kl_white_kernel<<<batch, 1024, 0, st>>>(q_mu, q_sqrt, M, kl, dmu, dLq, tril_out);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ overlap-add
// merged_mean / merged_variance of gpitch/window_overlap.py:19-59 on the device: per-window predictions [W, ws]
// (50 % overlap, hop = (ws-1)/2) -> one stream [n].  `win` is scipy's symmetric hann(ws) (power 1) or its square
// (power 2), supplied by the host so the weights are bit-identical; first window's left half and last window's right
// half have weight 1; products and the single add are rounded separately (no FMA) -> bit-exact vs the host code.
__global__ void __launch_bounds__(256) overlap_add_kernel(const double* __restrict__ Y, const double* __restrict__ win,
                                                          int nw, int ws, int n, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int half = (ws - 1) / 2;
  auto wgt = [&](int w, int k) -> double {               // weight of sample k of window w
    if (w == 0 && k < half) return 1.0;
    if (w == nw - 1 && nw > 1 && k >= ws - half) return 1.0;
    return win[k];
  };
  double v = 0.0;
  if (j < half) {
    v = __dmul_rn(Y[j], wgt(0, j));
  } else if (j >= n - half) {
    const int k = ws - (n - j);
    v = __dmul_rn(Y[(long long)(nw - 1) * ws + k], wgt(nw - 1, k));
  } else if (nw > 1 && j <= nw * half) {
    int i, o;
    if (j == nw * half) { i = nw - 2; o = half; }        // closing sample of the last loop iteration
    else { i = j / half - 1; o = j - (i + 1) * half; }
    const double a = __dmul_rn(Y[(long long)i * ws + half + o], wgt(i, half + o));
    const double b = __dmul_rn(Y[(long long)(i + 1) * ws + o], wgt(i + 1, o));
    v = __dadd_rn(a, b);
  }
  out[j] = v;
}

int launch_overlap_add(const double* Y, const double* win, int nw, int ws, int n, double* out, cudaStream_t st) {
  if (n <= 0) return GPX_OK;
  if (nw < 1 || ws < 3) return GPX_ERR_ARG;
  overlap_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(Y, win, nw, ws, n, out);
  GPX_CHECK_LAUNCH();
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ on-grid inducing points
// With z_j = x[iz_j] (inducing points on the window's sample grid) the M x M matrix K(z, z) is a column gather of the
// M x N matrix K(z, x): Kuu[m, j] = Kuf[m, iz_j] (+ jitter on the diagonal) -- the same kernel function of the same fp64
// inputs, so no second builder launch -- and its adjoint is a column scatter: Kuf_bar[m, iz_j] += Kuu_bar[m, j], after which
// ONE gradient pass over Kuf_bar yields the hyper-gradient of both matrices (gpitch/sgpr_ss.py:42-43 builds the two with the
// same kernel object).  iz_j < 0 marks a pad point of a ragged inducing set (init_models.pad_inducing): its row and column
// of Kuu are e_j * pad_diag, and it receives no adjoint (its true weight is exp(-1e3 / l) = 0).
__global__ void __launch_bounds__(256) gather_cols_kernel(const double* __restrict__ Kuf, long long sF, int ldf,
                                                          const int* __restrict__ iz, int div, int M,
                                                          const double* __restrict__ pad_diag, double jitter,
                                                          double* __restrict__ Kuu) {
  const int b = blockIdx.y;
  const int* izr = iz + (long long)(b / div) * M;
  const double* F = Kuf + (long long)b * sF;
  double* U = Kuu + (long long)b * M * M;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * M; e += gridDim.x * blockDim.x) {
    const int m = e / M, j = e - m * M;
    const int c = izr[j], cm = izr[m];
    double v;
    if (c < 0 || cm < 0) v = (m == j) ? pad_diag[b] : 0.0;
    else v = F[(long long)m * ldf + c];
    if (m == j) v += jitter;
    U[e] = v;
  }
}
__global__ void __launch_bounds__(256) scatter_add_cols_kernel(const double* __restrict__ Kuu_bar,
                                                               const int* __restrict__ iz, int div, int M,
                                                               double* __restrict__ Kuf_bar, long long sF, int ldf) {
  const int b = blockIdx.y;
  const int* izr = iz + (long long)(b / div) * M;
  const double* U = Kuu_bar + (long long)b * M * M;
  double* F = Kuf_bar + (long long)b * sF;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * M; e += gridDim.x * blockDim.x) {
    const int m = e / M, j = e - m * M;
    const int c = izr[j];
    if (c >= 0 && izr[m] >= 0) F[(long long)m * ldf + c] += U[e];     // distinct j -> distinct columns: no atomics
  }
}

int launch_gather_cols(const double* Kuf, long long sF, int ldf, const int* iz, int div, int M, const double* pad_diag,
                       double jitter, double* Kuu, int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    if (b0 % div) return GPX_ERR_ARG;
    const int gx = (M * M + 255) / 256 < 64 ? (M * M + 255) / 256 : 64;
    gather_cols_kernel<<<dim3(gx, nb), 256, 0, st>>>(Kuf + (long long)b0 * sF, sF, ldf, iz + (long long)(b0 / div) * M, div, M,
                                                     pad_diag + b0, jitter, Kuu + (long long)b0 * M * M);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}
int launch_scatter_add_cols(const double* Kuu_bar, const int* iz, int div, int M, double* Kuf_bar, long long sF, int ldf,
                            int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    if (b0 % div) return GPX_ERR_ARG;
    const int gx = (M * M + 255) / 256 < 64 ? (M * M + 255) / 256 : 64;
    scatter_add_cols_kernel<<<dim3(gx, nb), 256, 0, st>>>(Kuu_bar + (long long)b0 * M * M, iz + (long long)(b0 / div) * M, div, M,
                                                          Kuf_bar + (long long)b0 * sF, sF, ldf);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ packed lower triangles
// q_sqrt / dLq travel between host and device as packed lower triangles (row-major: element (i, j <= i) at
// i (i + 1) / 2 + j): the reference keeps q_sqrt as a dense M x M Param but only ever reads tf.matrix_band_part(., -1, 0)
// (GPflow conditional / gauss_kl, gpitch/pdgp.py:120-155), so the strict upper triangle carries no information and
// its gradient is identically zero.  Halves the PCIe bytes of the host-facing path.
__global__ void __launch_bounds__(256) tril_unpack_kernel(const double* __restrict__ packed, double* __restrict__ dense,
                                                          int M, long long T) {
  const long long b = blockIdx.y;
  const int i = blockIdx.x;                          // one CTA per row
  const double* src = packed + b * T + (long long)i * (i + 1) / 2;
  double* dst = dense + (b * M + i) * (long long)M;
  for (int j = threadIdx.x; j < M; j += blockDim.x) dst[j] = (j <= i) ? src[j] : 0.0;
}
__global__ void __launch_bounds__(256) tril_pack_kernel(const double* __restrict__ dense, double* __restrict__ packed,
                                                        int M, long long T) {
  const long long b = blockIdx.y;
  const int i = blockIdx.x;
  const double* src = dense + (b * M + i) * (long long)M;
  double* dst = packed + b * T + (long long)i * (i + 1) / 2;
  for (int j = threadIdx.x; j <= i; j += blockDim.x) dst[j] = src[j];
}

int launch_tril_unpack(const double* packed, double* dense, int M, int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  const long long T = (long long)M * (M + 1) / 2;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    tril_unpack_kernel<<<dim3(M, nb), 128, 0, st>>>(packed + b0 * T, dense + (long long)b0 * M * M, M, T);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}
int launch_tril_pack(const double* dense, double* packed, int M, int batch, cudaStream_t st) {
  if (batch <= 0 || M <= 0) return GPX_OK;
  const long long T = (long long)M * (M + 1) / 2;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    tril_pack_kernel<<<dim3(M, nb), 128, 0, st>>>(dense + (long long)b0 * M * M, packed + b0 * T, M, T);
    GPX_CHECK_LAUNCH();
  }
  return GPX_OK;
}

// ------------------------------------------------------------------------------------------ FP64 pipe peak
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const double A = a + threadIdx.x * 1e-9, B = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], A, B);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

int dmma_peak(int reps, double* tflops, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return GPX_ERR_LAUNCH;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = sms * 4;
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0, st);
    dmma_peak_kernel<<<grid, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return GPX_ERR_LAUNCH;
  *tflops = (double)grid * 8 * iters * 8.0 * 512 / (best * 1e-3) * 1e-12;
  return GPX_OK;
}

}  // namespace gpx

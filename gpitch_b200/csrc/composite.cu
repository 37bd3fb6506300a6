// Composite entry point: the collapsed (Titsias) bound of SGPRSS.build_likelihood (gpitch/sgpr_ss.py:29-71) and its
// gradient w.r.t. every kernel hyper-parameter and the noise variance, for W windows, as ONE C call -- the launch sequence
// of gpitch_b200/functions.py:SGPRBound + KernelMatrix / KernelPairOnGrid, so that a host without torch (the ctypes stub of
// INTEGRATION.md) can evaluate what GPflow's Model._objective evaluates.  Everything is stream-ordered; all scratch comes
// from the caller (gpx_sgpr_bound_workspace_bytes).
//
//   Kuf = sum_p k_p(z, x), Kuu = sum_p k_p(z, z) + jitter I, L = chol(Kuu), A = L^-1 Kuf / sigma, B = I + A A^T, LB = chol(B),
//   c = LB^-1 A y / sigma,
//   bound = -N/2 log 2pi - sum log diag(LB) - N/2 log s2 - y.y / (2 s2) + c.c / 2 - N sum_p kdiag_p / (2 s2) + tr(A A^T) / 2
//           [- 1000 sum_p |variance_p|  if reg]                                                     (sgpr_ss.py:40-68)
//   backward: v = LB^-T c, Abar = (I - B^-1) A + v w^T, w = y / sigma - A^T v,
//   Kuf_bar = L^-T Abar / sigma, Kuu_bar = -1/2 L^-T S L^-1 with S = B - 2 I + B^-1 + v v^T, then the kernel-gradient pass(es).
#include "../../include/gpitch_b200.h"
#include "builder.cuh"
#include "chol.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace gpx {
namespace {

constexpr double LOG2PI_C = 1.8378770664093453;

// per window: 1 / sigma, y.y, N * sum_p kdiag_p
__global__ void __launch_bounds__(256) sgpr_prep_kernel(const double* __restrict__ y, const double* __restrict__ noise,
                                                        const double* __restrict__ hyp, int kind, int N, int P, int Q,
                                                        double* __restrict__ inv_sigma, double* __restrict__ yy,
                                                        double* __restrict__ skd) {
  __shared__ double red[32];
  const int w = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const double v = y[(long long)w * N + i]; s = fma(v, v, s); }
  s = block_sum<false>(s, red);
  if (threadIdx.x == 0) {
    const int HS = 2 + 2 * Q;
    double kd = 0.0;
    for (int p = 0; p < P; p++) {
      const double* h = hyp + ((long long)w * P + p) * HS;
      double es = 1.0;
      if (kind != KIND_MATERN32) { es = 0.0; for (int q = 0; q < Q; q++) es += h[2 + q]; }
      kd += h[0] * es;
    }
    inv_sigma[w] = rsqrt(noise[w]);
    yy[w] = s;
    skd[w] = kd * N;
  }
}

// tr[w] = sum_i B[i][i]; then B[i][i] += add
__global__ void __launch_bounds__(256) diag_trace_add_kernel(double* __restrict__ B, int M, double add, double* __restrict__ tr) {
  __shared__ double red[32];
  double* Bw = B + (long long)blockIdx.x * M * M;
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double d = Bw[(long long)i * M + i];
    s += d;
    Bw[(long long)i * M + i] = d + add;
  }
  s = block_sum<false>(s, red);
  if (threadIdx.x == 0 && tr) tr[blockIdx.x] = s;
}

// c *= inv_sigma; bound from its pieces
__global__ void __launch_bounds__(256) sgpr_bound_kernel(double* __restrict__ c, const double* __restrict__ LB, int M, int N,
                                                         const double* __restrict__ inv_sigma, const double* __restrict__ noise,
                                                         const double* __restrict__ yy, const double* __restrict__ skd,
                                                         const double* __restrict__ trAAT, const double* __restrict__ hyp,
                                                         int P, int HS, int reg, double* __restrict__ bound) {
  __shared__ double red[32];
  const int w = blockIdx.x;
  double cc = 0.0, ld = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double v = c[(long long)w * M + i] * inv_sigma[w];
    c[(long long)w * M + i] = v;
    cc = fma(v, v, cc);
    ld += log(LB[((long long)w * M + i) * M + i]);
  }
  cc = block_sum<false>(cc, red);
  ld = block_sum<false>(ld, red);
  if (threadIdx.x == 0) {
    const double s2 = noise[w];
    double b = -0.5 * N * LOG2PI_C - ld - 0.5 * N * log(s2) - 0.5 * yy[w] / s2 + 0.5 * cc - 0.5 * skd[w] / s2 + 0.5 * trAAT[w];
    if (reg) {
      double r = 0.0;
      for (int p = 0; p < P; p++) r += fabs(hyp[((long long)w * P + p) * HS]);
      b -= 1000.0 * r;
    }
    bound[w] = b;
  }
}

// wv[n] = y[n] / sigma - Atv[n];  uAtv[w] = sum_n (y[n] / sigma) Atv[n]
__global__ void __launch_bounds__(256) sgpr_w_kernel(const double* __restrict__ y, const double* __restrict__ Atv,
                                                     const double* __restrict__ inv_sigma, int N, double* __restrict__ wv,
                                                     double* __restrict__ uAtv) {
  __shared__ double red[32];
  const int w = blockIdx.x;
  const double is = inv_sigma[w];
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double u = y[(long long)w * N + i] * is, a = Atv[(long long)w * N + i];
    wv[(long long)w * N + i] = u - a;
    s = fma(u, a, s);
  }
  s = block_sum<false>(s, red);
  if (threadIdx.x == 0) uAtv[w] = s;
}

// ImB = I - Binv;  S = B + Binv + v v^T - 2 I (B already holds A A^T + I);  trS[w]
__global__ void __launch_bounds__(256) sgpr_imb_s_kernel(const double* __restrict__ B, const double* __restrict__ Binv,
                                                         const double* __restrict__ v, int M, double* __restrict__ ImB,
                                                         double* __restrict__ S, double* __restrict__ trS) {
  __shared__ double red[32];
  const int w = blockIdx.y;
  const long long o = (long long)w * M * M;
  double t = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * M; e += gridDim.x * blockDim.x) {
    const int i = e / M, j = e - i * M;
    const double bi = Binv[o + e], d = (i == j) ? 1.0 : 0.0;
    ImB[o + e] = d - bi;
    const double s = B[o + e] + bi + v[(long long)w * M + i] * v[(long long)w * M + j] - 2.0 * d;
    S[o + e] = s;
    if (i == j) t += s;
  }
  t = block_sum<false>(t, red);
  if (threadIdx.x == 0) atomicAdd(trS + w, t);
}

__global__ void scale_rows_kernel(double* __restrict__ v, const double* __restrict__ s, int M) {   // v[w, :] *= s[w]
  const int w = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) v[(long long)w * M + i] *= s[w];
}

// dhyp = dhyp_uf (+ dhyp_uu) + the kdiag and regulariser terms;  dnoise
__global__ void sgpr_grad_final_kernel(const double* __restrict__ hyp, const double* __restrict__ noise, int kind, int N,
                                       int P, int Q, int reg, const double* __restrict__ d2, const double* __restrict__ yy,
                                       const double* __restrict__ skd, const double* __restrict__ trS,
                                       const double* __restrict__ uAtv, double* __restrict__ dhyp, double* __restrict__ dnoise) {
  const int w = blockIdx.x, HS = 2 + 2 * Q;
  const double s2 = noise[w];
  for (int e = threadIdx.x; e < P * HS; e += blockDim.x) {
    const int p = e / HS, k = e - p * HS;
    const double* h = hyp + ((long long)w * P + p) * HS;
    double g = dhyp[(long long)w * P * HS + e] + (d2 ? d2[(long long)w * P * HS + e] : 0.0);
    if (k == 0) {                                  // d(-N kdiag / (2 s2)) / d variance, regulariser
      double es = 1.0;
      if (kind != KIND_MATERN32) { es = 0.0; for (int q = 0; q < Q; q++) es += h[2 + q]; }
      g += -0.5 * N * es / s2;
      if (reg) g -= 1000.0 * (h[0] > 0.0 ? 1.0 : (h[0] < 0.0 ? -1.0 : 0.0));
    } else if (k >= 2 && k < 2 + Q && kind != KIND_MATERN32) {
      g += -0.5 * N * h[0] / s2;                   // d / d energy_q
    }
    dhyp[(long long)w * P * HS + e] = g;
  }
  if (threadIdx.x == 0 && dnoise)
    dnoise[w] = -0.5 * N / s2 + 0.5 * yy[w] / (s2 * s2) + 0.5 * skd[w] / (s2 * s2) - (trS[w] + uAtv[w]) / (2.0 * s2);
}

GemmArgs gargs(int batch, const double* A, long long sA, int lda, const double* B, long long sB, int ldb, double* C,
               long long sC, int ldc, int M, int N, int K, int flags) {
  GemmArgs g = {};
  g.A = A; g.B = B; g.C = C; g.sA = sA; g.sB = sB; g.sC = sC; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.batch = batch; g.flags = flags; g.alpha = 1.0;
  return g;
}

struct Ws {      // workspace carving (doubles)
  double* p;
  double* take(long long n) { double* r = p; p += (n + 1) & ~1LL; return r; }   // keep 16-byte alignment
};

long long ws_doubles(int kind, int N, int M, int P, int Q, int W, int with_grad, int nlag) {
  const long long KP = kind == KIND_MERCER_M12 ? feat_rows(Q) : 0, HS = 2 + 2 * Q;
  long long n = 0;
  auto add = [&](long long k) { n += (k + 1) & ~1LL; };
  add((long long)W * P * KP * M); add((long long)W * P * KP * N);          // features
  const long long Np = (N + 1) & ~1LL;                                     // row pitch of the M x N matrices (see gpx_sgpr_bound)
  add((long long)W * M * Np); add((long long)W * M * Np);                  // Kuf, A
  for (int i = 0; i < 6; i++) add((long long)W * M * M);                   // Kuu/L, Linv, B, LB, LBinv, scratch
  add((long long)W * 64 * M);                                              // potrf work
  for (int i = 0; i < 8; i++) add(W);                                      // per-window scalars
  add((long long)W * M); add((long long)W * M);                            // Aerr, c
  if (with_grad) {
    add((long long)W * M * Np);                                            // Kuf_bar
    for (int i = 0; i < 4; i++) add((long long)W * M * M);                 // Binv, ImB / U, H, S
    add((long long)W * M); add((long long)W * M);                          // v, av
    add((long long)W * N); add((long long)W * N);                          // Atv, w
    add((long long)W * (N > M ? N : M)); add(W);                           // dummy fvar rows, dummy kdiag
    add((long long)W * P * HS);                                            // second hyper-gradient
    if (nlag > 0) add((long long)W * P * ((long long)N + 2LL * nlag));     // lag-histogram scratch
  }
  return n;
}

}  // namespace
}  // namespace gpx

extern "C" {

long long gpx_sgpr_bound_workspace_bytes(int kind, int N, int M, int P, int Q, int W, int with_grad, int nlag) {
  if (N < 1 || M < 1 || P < 1 || Q < 0 || W < 0 || kind < 0 || kind > 3) return -1;
  return 8 * gpx::ws_doubles(kind, N, M, P, Q, W, with_grad, nlag);
}

int gpx_sgpr_bound(int kind, int mode, const double* x, const double* y, const double* z, int N, int M, int W,
                   const double* hyp, int P, int Q, const double* noise, double jitter, int reg, int need_ef, const int* iz,
                   const double* delta, int nlag, double* bound, double* dhyp, double* dnoise, int* info, double* work,
                   void* stream) {
  using namespace gpx;
  if (W <= 0) return GPX_OK;
  if (!x || !y || !z || !hyp || !noise || !bound || !info || !work || N < 1 || M < 1 || P < 1 || W > 65535 ||
      kind < 0 || kind > 3 || (kind == KIND_MERCER_M12 && Q < 1))
    return GPX_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = dhyp != nullptr;
  const bool lag = iz && delta && nlag >= N && kind == KIND_MERCER_M12;      // inducing points on the sample grid
  const int HS = 2 + 2 * Q, KP = kind == KIND_MERCER_M12 ? feat_rows(Q) : 0;
  // The M x N matrices (Kuf, A, Kuf_bar) get an even row pitch: TMA needs 16-byte-aligned rows, and gpitch's usual window
  // (ws = 2001 samples) is odd -- without the pad column every GEMM that touches them falls back to the cp.async kernel.
  const int Np = (N + 1) & ~1;
  const long long MM = (long long)M * M, MN = (long long)M * Np;
  Ws ws{work};
  double* fz = ws.take((long long)W * P * KP * M);
  double* fx = ws.take((long long)W * P * KP * N);
  double* Kuf = ws.take(W * MN);
  double* A = ws.take(W * MN);
  double* Kuu = ws.take(W * MM);     // becomes L
  double* Linv = ws.take(W * MM);
  double* B = ws.take(W * MM);
  double* LB = ws.take(W * MM);
  double* LBinv = ws.take(W * MM);
  double* T1 = ws.take(W * MM);
  double* pwork = ws.take((long long)W * 64 * M);
  double* inv_sigma = ws.take(W); double* yy = ws.take(W); double* skd = ws.take(W); double* trAAT = ws.take(W);
  double* trS = ws.take(W); double* uAtv = ws.take(W); double* scale = ws.take(W); double* zeros = ws.take(W);
  double* Aerr = ws.take((long long)W * M);
  double* c = ws.take((long long)W * M);
  int rc;
#define RUN(call) do { if ((rc = (call)) != GPX_OK) return rc; } while (0)

  // ---- forward
  sgpr_prep_kernel<<<W, 256, 0, st>>>(y, noise, hyp, kind, N, P, Q, inv_sigma, yy, skd);
  GPX_CHECK_LAUNCH();
  if (kind == KIND_MERCER_M12) {
    RUN(launch_features(z, M, 1, hyp, P, Q, fz, W, st));
    RUN(launch_features(x, N, 1, hyp, P, Q, fx, W, st));
  }
  KernArgs k = {};
  k.kind = kind; k.mode = mode; k.ptsA = z; k.nA = M; k.divA = 1; k.hyp = hyp; k.P = P; k.Q = Q; k.batch = W;
  k.featA = KP ? fz : nullptr;
  KernArgs kf = k;
  kf.ptsB = x; kf.nB = N; kf.divB = 1; kf.featB = KP ? fx : nullptr; kf.K = Kuf; kf.sK = MN; kf.ldk = Np; kf.jitter = 0.0;
  RUN(launch_kernel_build(kf, st));
  KernArgs ku = k;
  ku.ptsB = z; ku.nB = M; ku.divB = 1; ku.featB = KP ? fz : nullptr; ku.K = Kuu; ku.sK = MM; ku.ldk = M; ku.jitter = jitter;
  if (lag) {          // K(z, z) = the columns iz of K(z, x): no second builder launch (pad points: decoupled diagonal)
    RUN(launch_gather_cols(Kuf, MN, Np, iz, 1, M, skd, jitter, Kuu, W, st));
  } else {
    RUN(launch_kernel_build(ku, st));
  }
  RUN(potrf_trinv(Kuu, MM, M, Linv, MM, M, pwork, info, M, W, st));
  {
    GemmArgs g = gargs(W, Linv, MM, M, Kuf, MN, Np, A, MN, Np, M, N, M, GEMM_A_LOWER);
    g.alpha_vec = inv_sigma;
    RUN(launch_gemm(g, st));
    GemmArgs s = gargs(W, A, MN, Np, A, MN, Np, B, MM, M, M, M, N, GEMM_TRANS_B | GEMM_C_LOWER | GEMM_C_MIRROR);
    RUN(launch_gemm(s, st));
  }
  diag_trace_add_kernel<<<W, 256, 0, st>>>(B, M, 1.0, trAAT);
  GPX_CHECK_LAUNCH();
  if (cudaMemcpyAsync(LB, B, sizeof(double) * W * MM, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return GPX_ERR_LAUNCH;
  RUN(potrf_trinv(LB, MM, M, LBinv, MM, M, pwork, info + W, M, W, st));
  RUN(launch_rowdot(A, MN, Np, y, N, Aerr, M, N, W, st));                       // A y
  RUN(launch_rowdot(LBinv, MM, M, Aerr, M, c, M, M, W, st));                   // LB^-1 (A y)
  sgpr_bound_kernel<<<W, 256, 0, st>>>(c, LB, M, N, inv_sigma, noise, yy, skd, trAAT, hyp, P, HS, reg, bound);
  GPX_CHECK_LAUNCH();
  if (!grad) return GPX_OK;

  // ---- backward
  double* Kufb = ws.take(W * MN);
  double* Binv = ws.take(W * MM);
  double* ImB = ws.take(W * MM);     // later reused as U
  double* H = ws.take(W * MM);
  double* S = ws.take(W * MM);
  double* v = ws.take((long long)W * M);
  double* av = ws.take((long long)W * M);
  double* Atv = ws.take((long long)W * N);
  double* wv = ws.take((long long)W * N);
  double* dummyN = ws.take((long long)W * (N > M ? N : M));
  double* dummyW = ws.take(W);
  double* dhyp2 = ws.take((long long)W * P * HS);
  double* lagwork = lag ? ws.take((long long)W * P * ((long long)N + 2LL * nlag)) : nullptr;
  cudaMemsetAsync(zeros, 0, sizeof(double) * W, st);
  cudaMemsetAsync(trS, 0, sizeof(double) * W, st);
  // v = LB^-T c  (column statistics of LB^-1 with mu = c: fmean[n] = sum_m LBinv[m, n] c[m])
  RUN(launch_cond_colstats(LBinv, nullptr, MM, M, c, zeros, v, dummyN, M, M, W, 0, st));
  {
    GemmArgs g = gargs(W, LBinv, MM, M, LBinv, MM, M, Binv, MM, M, M, M, M,
                       GEMM_TRANS_A | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_C_LOWER | GEMM_C_MIRROR);
    RUN(launch_gemm(g, st));
  }
  RUN(launch_cond_colstats(A, nullptr, MN, Np, v, zeros, Atv, dummyN, M, N, W, 0, st));      // A^T v
  sgpr_w_kernel<<<W, 256, 0, st>>>(y, Atv, inv_sigma, N, wv, uAtv);
  GPX_CHECK_LAUNCH();
  {
    const int gx = (int)((MM + 255) / 256 < 64 ? (MM + 255) / 256 : 64);
    sgpr_imb_s_kernel<<<dim3(gx, W), 256, 0, st>>>(B, Binv, v, M, ImB, S, trS);
    GPX_CHECK_LAUNCH();
  }
  {
    GemmArgs h = gargs(W, Linv, MM, M, ImB, MM, M, H, MM, M, M, M, M, GEMM_TRANS_A | GEMM_A_UPPER);   // L^-T (I - B^-1)
    RUN(launch_gemm(h, st));
  }
  RUN(launch_cond_colstats(Linv, nullptr, MM, M, v, zeros, av, dummyN, M, M, W, 0, st));               // L^-T v
  scale_rows_kernel<<<dim3((M + 255) / 256, W), 256, 0, st>>>(av, inv_sigma, M);
  GPX_CHECK_LAUNCH();
  {
    GemmArgs g = gargs(W, H, MM, M, A, MN, Np, Kufb, MN, Np, M, N, M, 0);        // Kuf_bar = H A / sigma + (L^-T v / sigma) w^T
    g.alpha_vec = inv_sigma; g.rowvec = av; g.colvec = wv; g.sRowvec = M; g.sColvec = N;
    RUN(launch_gemm(g, st));
    double* U = ImB;
    GemmArgs u = gargs(W, S, MM, M, Linv, MM, M, U, MM, M, M, M, M, GEMM_B_LOWER);
    RUN(launch_gemm(u, st));
    GemmArgs d = gargs(W, Linv, MM, M, U, MM, M, T1, MM, M, M, M, M, GEMM_TRANS_A | GEMM_A_UPPER | GEMM_C_LOWER | GEMM_C_MIRROR);
    d.alpha = -0.5;
    RUN(launch_gemm(d, st));                                                    // Kuu_bar in T1
  }
  KernArgs gf = kf;
  gf.K = Kufb; gf.dhyp = dhyp; gf.need_ef = need_ef;
  if (lag) {        // inducing points on the sample grid: scatter Kuu_bar into Kuf_bar, ONE lag-histogram pass
    RUN(launch_scatter_add_cols(T1, iz, 1, M, Kufb, MN, Np, W, st));
    RUN(launch_kernel_grad_lag(gf, iz, delta, lagwork, nlag, st));
    dhyp2 = nullptr;
  } else {
    cudaMemsetAsync(dhyp, 0, sizeof(double) * (size_t)W * P * HS, st);
    cudaMemsetAsync(dhyp2, 0, sizeof(double) * (size_t)W * P * HS, st);
    RUN(launch_kernel_grad(gf, st));
    KernArgs gu = ku;
    gu.K = T1; gu.dhyp = dhyp2; gu.need_ef = need_ef; gu.jitter = 0.0;
    RUN(launch_kernel_grad(gu, st));
  }
  sgpr_grad_final_kernel<<<W, 128, 0, st>>>(hyp, noise, kind, N, P, Q, reg, dhyp2, yy, skd, trS, uAtv, dhyp, dnoise);
  GPX_CHECK_LAUNCH();
  (void)dummyW; (void)scale;
#undef RUN
  return GPX_OK;
}

}  // extern "C"

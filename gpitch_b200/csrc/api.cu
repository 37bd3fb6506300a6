// extern "C" entry points declared in include/gpitch_b200.h.
#include "../../include/gpitch_b200.h"
#include "builder.cuh"
#include "chol.cuh"
#include "gemm.cuh"
#include "ops.cuh"
#include <cstring>

namespace gpx {
int set_hermgauss(const double* x, const double* w, int n);
int dmma_peak(int reps, double* tflops, cudaStream_t st);
}

static_assert(sizeof(gpx_gemm_args) == sizeof(gpx::GemmArgs), "ABI struct must mirror gpx::GemmArgs");

namespace gpx {
std::atomic<unsigned long long> g_launches{0};
extern std::atomic<unsigned long long> g_gemm_tma_launches;
}

extern "C" {

unsigned long long gpx_launch_count(void) { return gpx::g_launches.load(); }

unsigned long long gpx_gemm_tma_launch_count(void) { return gpx::g_gemm_tma_launches.load(); }

int gpx_version(void) { return 200; }

int gpx_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? GPX_OK : GPX_ERR_ARG; }

int gpx_set_hermgauss(const double* x, const double* w, int n) { return gpx::set_hermgauss(x, w, n); }

int gpx_feat_rows(int Q) { return gpx::feat_rows(Q); }

int gpx_features(const double* pts, int n, int div, const double* hyp, int P, int Q, double* feat, int batch,
                 void* stream) {
  if (!pts || !hyp || !feat || div < 1) return GPX_ERR_ARG;
  return gpx::launch_features(pts, n, div, hyp, P, Q, feat, batch, (cudaStream_t)stream);
}

static int fill_kern(gpx::KernArgs& a, int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB,
                     int nB, int divB, const double* hyp, int P, int Q, const double* featA, const double* featB,
                     double* K, long long strideK, int ldk, int batch) {
  if (kind < 0 || kind > 3 || mode < 0 || mode > 1 || !ptsA || !ptsB || !hyp || !K || divA < 1 || divB < 1 ||
      ldk < nB || Q < 0)
    return GPX_ERR_ARG;
  a = gpx::KernArgs{};
  a.kind = kind; a.mode = mode;
  a.ptsA = ptsA; a.ptsB = ptsB; a.nA = nA; a.nB = nB; a.divA = divA; a.divB = divB;
  a.hyp = hyp; a.P = P; a.Q = Q; a.featA = featA; a.featB = featB;
  a.K = K; a.sK = strideK; a.ldk = ldk; a.batch = batch;
  return GPX_OK;
}

int gpx_kernel_build(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB, int divB,
                     const double* hyp, int P, int Q, const double* featA, const double* featB, double* K,
                     long long strideK, int ldk, double jitter, int batch, void* stream) {
  gpx::KernArgs a;
  int rc = fill_kern(a, kind, mode, ptsA, nA, divA, ptsB, nB, divB, hyp, P, Q, featA, featB, K, strideK, ldk, batch);
  if (rc) return rc;
  a.jitter = jitter;
  return gpx::launch_kernel_build(a, (cudaStream_t)stream);
}

int gpx_kernel_grad(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB, int divB,
                    const double* hyp, int P, int Q, const double* featA, const double* featB, const double* Kbar,
                    long long strideK, int ldk, double* dhyp, int need_ef, const double* epi_col,
                    const double* epi_rowv, const double* epi_colv, double epi_alpha, double* dptsA, int batch,
                    void* stream) {
  gpx::KernArgs a;
  int rc = fill_kern(a, kind, mode, ptsA, nA, divA, ptsB, nB, divB, hyp, P, Q, featA, featB,
                     const_cast<double*>(Kbar), strideK, ldk, batch);
  if (rc) return rc;
  if (!dhyp) return GPX_ERR_ARG;
  a.dhyp = dhyp; a.need_ef = need_ef;
  a.epi_col = epi_col; a.epi_rowv = epi_rowv; a.epi_colv = epi_colv; a.epi_alpha = epi_alpha;
  a.dpts = dptsA;
  if (batch > 0)
    cudaMemsetAsync(dhyp, 0, sizeof(double) * (size_t)batch * P * (2 + 2 * Q), (cudaStream_t)stream);
  if (batch > 0 && dptsA) cudaMemsetAsync(dptsA, 0, sizeof(double) * (size_t)batch * nA, (cudaStream_t)stream);
  return gpx::launch_kernel_grad(a, (cudaStream_t)stream);
}

int gpx_kernel_grad_lag(int mode, const double* ptsA, int nA, int divA, const int* izA, const double* ptsB, int nB, int divB,
                        const double* delta, const double* hyp, int P, int Q, const double* Kbar, long long strideK,
                        int ldk, double* dhyp, int need_ef, const double* epi_col, const double* epi_rowv,
                        const double* epi_colv, double epi_alpha, double* work, int nlag, int batch, void* stream) {
  gpx::KernArgs a;
  int rc = fill_kern(a, GPX_KIND_MERCER_M12, mode, ptsA, nA, divA, ptsB, nB, divB, hyp, P, Q, nullptr, nullptr,
                     const_cast<double*>(Kbar), strideK, ldk, batch);
  if (rc) return rc;
  if (!dhyp || !izA || !delta || !work) return GPX_ERR_ARG;
  a.dhyp = dhyp; a.need_ef = need_ef;
  a.epi_col = epi_col; a.epi_rowv = epi_rowv; a.epi_colv = epi_colv; a.epi_alpha = epi_alpha;
  return gpx::launch_kernel_grad_lag(a, izA, delta, work, nlag, (cudaStream_t)stream);
}

int gpx_kernel_grad_points(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB, int divB,
                           const double* hyp, int P, int Q, const double* featA, const double* featB, const double* Kbar,
                           long long strideK, int ldk, double* dptsA, const double* epi_col, const double* epi_rowv,
                           const double* epi_colv, double epi_alpha, int batch, void* stream) {
  gpx::KernArgs a;
  int rc = fill_kern(a, kind, mode, ptsA, nA, divA, ptsB, nB, divB, hyp, P, Q, featA, featB,
                     const_cast<double*>(Kbar), strideK, ldk, batch);
  if (rc) return rc;
  a.epi_col = epi_col; a.epi_rowv = epi_rowv; a.epi_colv = epi_colv; a.epi_alpha = epi_alpha;
  return gpx::launch_kernel_grad_points(a, dptsA, (cudaStream_t)stream);
}

long long gpx_potrf_workspace_bytes(int M, int batch) {
  return (M < 1 || batch < 0) ? -1 : (long long)sizeof(double) * batch * 64 * M;
}

long long gpx_kernel_grad_lag_workspace_bytes(int nB, int P, int nlag, int batch) {
  return (nB < 1 || P < 1 || nlag < nB || batch < 0) ? -1 : (long long)sizeof(double) * batch * P * ((long long)nB + 2LL * nlag);
}

int gpx_potrf_trinv(double* A, long long strideA, int lda, double* Linv, long long strideI, int ldi, double* work,
                    int* info, int M, int batch, void* stream) {
  if (lda < M || ldi < M) return GPX_ERR_ARG;
  return gpx::potrf_trinv(A, strideA, lda, Linv, strideI, ldi, work, info, M, batch, (cudaStream_t)stream);
}

int gpx_gemm(const gpx_gemm_args* args, void* stream) {
  if (!args || !args->A || !args->B || !args->C) return GPX_ERR_ARG;
  gpx::GemmArgs g;
  memcpy(&g, args, sizeof(g));
  if (g.rowvec && !g.colvec) return GPX_ERR_ARG;
  if (g.Aux && g.ldaux < g.N) return GPX_ERR_ARG;
  return gpx::launch_gemm(g, (cudaStream_t)stream);
}

int gpx_cond_colstats(const double* A, const double* LTA, long long strideA, int ld, const double* q_mu,
                      const double* kdiag, double* fmean, double* fvar, int M, int N, int batch, int mode,
                      void* stream) {
  if (!A || !q_mu || !kdiag || !fmean || !fvar || ld < N || mode < 0 || mode > 1 || (mode == 1 && !LTA)) return GPX_ERR_ARG;
  return gpx::launch_cond_colstats(A, LTA, strideA, ld, q_mu, kdiag, fmean, fvar, M, N, batch, mode, (cudaStream_t)stream);
}

int gpx_scale_rank1(const double* T, long long strideT, int ld, const double* colscale, const double* rowvec,
                    const double* colvec, double alpha, double* out, int M, int N, int batch, void* stream) {
  if (!T || !colscale || !rowvec || !colvec || !out || ld < N) return GPX_ERR_ARG;
  return gpx::launch_scale_rank1(T, strideT, ld, colscale, rowvec, colvec, alpha, out, M, N, batch, (cudaStream_t)stream);
}

int gpx_rowdot(const double* A, long long strideA, int ld, const double* v, long long strideV, double* out, int M,
               int N, int batch, void* stream) {
  if (!A || !v || !out || ld < N) return GPX_ERR_ARG;
  return gpx::launch_rowdot(A, strideA, ld, v, strideV, out, M, N, batch, (cudaStream_t)stream);
}

int gpx_varexp(const double* Fmu, const double* Fvar, const double* Y, const double* noise, int P, int W, int N,
               int nlin, double* ve_sum, double* dFmu, double* dFvar, double* dnoise, double* ve_pointwise,
               void* stream) {
  if (!Fmu || !Fvar || !Y || !noise || !ve_sum) return GPX_ERR_ARG;
  if ((dFmu == nullptr) != (dFvar == nullptr)) return GPX_ERR_ARG;
  if (W > 0) {
    cudaMemsetAsync(ve_sum, 0, sizeof(double) * (size_t)W, (cudaStream_t)stream);
    if (dnoise) cudaMemsetAsync(dnoise, 0, sizeof(double) * (size_t)W, (cudaStream_t)stream);
  }
  return gpx::launch_varexp(Fmu, Fvar, Y, noise, P, W, N, nlin, ve_sum, dFmu, dFvar, dnoise, ve_pointwise,
                            (cudaStream_t)stream);
}

int gpx_overlap_add(const double* Y, const double* win, int num_windows, int ws, int n, double* out, void* stream) {
  if (!Y || !win || !out) return GPX_ERR_ARG;
  return gpx::launch_overlap_add(Y, win, num_windows, ws, n, out, (cudaStream_t)stream);
}

int gpx_kuu_from_kuf(const double* Kuf, long long strideF, int ldf, const int* iz, int div, int M, const double* pad_diag,
                     double jitter, double* Kuu, int batch, void* stream) {
  if (!Kuf || !iz || !Kuu || !pad_diag || div < 1 || M < 1) return GPX_ERR_ARG;
  return gpx::launch_gather_cols(Kuf, strideF, ldf, iz, div, M, pad_diag, jitter, Kuu, batch, (cudaStream_t)stream);
}

int gpx_kuu_bar_into_kuf_bar(const double* Kuu_bar, const int* iz, int div, int M, double* Kuf_bar, long long strideF, int ldf,
                             int batch, void* stream) {
  if (!Kuu_bar || !iz || !Kuf_bar || div < 1 || M < 1) return GPX_ERR_ARG;
  return gpx::launch_scatter_add_cols(Kuu_bar, iz, div, M, Kuf_bar, strideF, ldf, batch, (cudaStream_t)stream);
}

int gpx_tril_unpack(const double* packed, double* dense, int M, int batch, void* stream) {
  if (!packed || !dense || M < 1) return GPX_ERR_ARG;
  return gpx::launch_tril_unpack(packed, dense, M, batch, (cudaStream_t)stream);
}

int gpx_tril_pack(const double* dense, double* packed, int M, int batch, void* stream) {
  if (!packed || !dense || M < 1) return GPX_ERR_ARG;
  return gpx::launch_tril_pack(dense, packed, M, batch, (cudaStream_t)stream);
}

int gpx_gauss_kl_white(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* dmu,
                       double* dLq, void* stream) {
  if (!q_mu || !q_sqrt || !kl || M < 1) return GPX_ERR_ARG;
  return gpx::launch_gauss_kl_white(q_mu, q_sqrt, M, batch, kl, dmu, dLq, (cudaStream_t)stream);
}


int gpx_gauss_kl_white_tril(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* tril_out,
                            void* stream) {
  if (!q_mu || !q_sqrt || !kl || !tril_out || M < 1) return GPX_ERR_ARG;
  return gpx::launch_gauss_kl_white(q_mu, q_sqrt, M, batch, kl, nullptr, nullptr, (cudaStream_t)stream, tril_out);
}

/* FP64 tensor-pipe peak: register-resident mma.sync m8n8k4 loop, best of `reps`; returns TFLOP/s in *tflops (host). */
int gpx_dmma_peak(int reps, double* tflops, void* stream) {
  return gpx::dmma_peak(reps, tflops, (cudaStream_t)stream);
}

}  // extern "C"

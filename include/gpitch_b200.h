/* gpitch_b200 -- C ABI of the B200-native variational-GP inner loop of gpitch.
 *
 * This is the drop-in boundary.  The reference (PabloAlvarado/gpitch) has no FFI layer of its own: its hot path
 * is a TensorFlow-1 graph built by GPflow-0.5 subclasses, so every entry point below cites the reference graph
 * stage (file:line under the reference tree) whose stock TF ops it replaces.  Host bindings: gpitch_b200/_lib.py
 * (ctypes); INTEGRATION.md shows the stub a gpitch maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (inputs, outputs and
 *    workspace; the two entry points that need scratch have a gpx_*_workspace_bytes() query); the compute entry points
 *    allocate nothing, retain nothing and never synchronise.  The only exceptions are one-time, per-device table uploads
 *    (gpx_set_hermgauss, the exp2 table on the first kernel-matrix call) and the measurement helper gpx_dmma_peak, which
 *    allocates a scratch buffer and synchronises -- it exists for bench.py, not for the data path;
 *  - thread safety: entry points may be called from several host threads on different streams; launch counters are
 *    atomic, per-device state is keyed by the current device;
 *  - fp64, row-major, contiguous last dimension, leading dimensions in elements; the window / latent-GP batch is
 *    the outermost dimension ("batch");
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered;
 *  - return value: 0 ok, -1 bad argument, -2 CUDA launch failure.  Numerical failure (non-PD matrix) is reported
 *    per batch entry in the device array `info` (LAPACK convention), never by the return code.
 */
#ifndef GPITCH_B200_H
#define GPITCH_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define GPX_KIND_MERCER_M12 0 /* MercerMatern12sm, gpitch/matern12_spectral_mixture.py:70-133 */
#define GPX_KIND_DIFF_M12 1   /* Matern12sm,       gpitch/matern12_spectral_mixture.py:14-67  */
#define GPX_KIND_MATERN32 2   /* gpflow.kernels.Matern32 (gpitch/init_kernels.py:12)          */
#define GPX_KIND_DIFF_M32 3   /* Matern32sm / Matern32sml (legacy), gpitch/kernels.py:204-318: variance = 1, energies = the
                                 per-partial variances; Matern32sml = one single-partial component per partial */
#define GPX_DIST_REFERENCE 0  /* GPflow Stationary.square_dist operation order (bit-reproducible) */
#define GPX_DIST_STABLE 1     /* direct |x - x'| (optional, better conditioned at large t)     */
#define GPX_NLIN_LOGISTIC 0   /* logistic_tf, gpitch/methods.py:216-218 */
#define GPX_NLIN_SOFTPLUS 1   /* softplus_tf, gpitch/methods.py:220-222 */
#define GPX_NLIN_GAUSS 2      /* gaussfun_tf, gpitch/methods.py:232-233 */

/* flags of gpx_gemm */
#define GPX_GEMM_TRANS_A 1
#define GPX_GEMM_TRANS_B 2
#define GPX_GEMM_A_LOWER 4
#define GPX_GEMM_A_UPPER 8
#define GPX_GEMM_B_LOWER 16
#define GPX_GEMM_B_UPPER 32
#define GPX_GEMM_C_LOWER 64
#define GPX_GEMM_C_MIRROR 128
#define GPX_GEMM_ZERO_UPPER 256

int gpx_version(void);

/* Bind the library's CUDA runtime to `device` for the calling thread (one process per GPU: call once with
 * LOCAL_RANK before any other entry point). */
int gpx_set_device(int device);

/* Gauss-Hermite nodes / weights (weights already divided by sqrt(pi)); replaces gpflow.quadrature.hermgauss(H)
 * as used by gpitch/likelihoods.py:35-37.  Host pointers.  Must be called once per process before gpx_varexp. */
int gpx_set_hermgauss(const double* x_host, const double* w_host, int n);

/* Mercer features phi = [sqrt(e_q) cos(2 pi f_q x); sqrt(e_q) sin(2 pi f_q x)] -- MercerMatern12sm.phi_features,
 * gpitch/matern12_spectral_mixture.py:123-133.
 *   pts  [batch/div, n]    point sets; batch entry b uses row b / div (div consecutive latent GPs share a window)
 *   hyp  [batch, P, 2+2Q]  (variance, lengthscale, energy[Q], frequency[Q]) per component kernel
 *   feat [batch, P, KP, n] out, KP = gpx_feat_rows(Q) (2Q rounded up to a multiple of 4, zero padded) */
int gpx_feat_rows(int Q);
int gpx_features(const double* pts, int n, int div, const double* hyp, int P, int Q, double* feat, int batch,
                 void* stream);

/* Fused covariance builder  K[b] = sum_p k_p(ptsA_b, ptsB_b) (+ jitter on the diagonal).
 * Replaces Kern.K of MercerMatern12sm (:102-117), Matern12sm (:38-56), GPflow Matern32 and the GPflow `Add`
 * of P pitch kernels (gpitch/transcription.py:245, gpitch/separation.py:257); call sites gpitch/sgpr_ss.py:42-43,
 * :88,:93 and GPflow conditional() from gpitch/pdgp.py:147-155.
 *   ptsA [batch/divA, nA], ptsB [batch/divB, nB]: batch entry b uses rows b / divA, b / divB
 *   K [batch, nA, ldk] out (batch stride strideK elements) */
int gpx_kernel_build(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB, int divB,
                     const double* hyp, int P, int Q, const double* featA, const double* featB, double* K,
                     long long strideK, int ldk, double jitter, int batch, void* stream);

/* Analytic hyper-parameter gradient  dhyp[b,p,:] = sum_mn Kbar[b,m,n] dK_p[m,n]/d(var, len, e_q, f_q).
 * Replaces tf.gradients through the builder graph (GPflow Model._objective; call sites gpitch/separation.py:298,
 * gpitch/transcription.py:283).  dhyp is overwritten.  need_ef = 0 skips energy/frequency (fixed params).
 * Optional fused epilogue on the incoming adjoint (epi_col = NULL: none), so that conditional()'s
 * Kbar_mn = 2 T diag(vbar) + a mbar^T is consumed straight from T without being written to memory:
 *   Kbar_eff[b,m,n] = epi_alpha * epi_col[b,n] * Kbar[b,m,n] + epi_rowv[b,m] * epi_colv[b,n]
 *   epi_col [batch, nB], epi_rowv [batch, nA] (NULL = 0), epi_colv [batch, nB] (NULL = 0)
 * dptsA (NULL = skip): [batch, nA] out, overwritten -- the row-point gradient of gpx_kernel_grad_points from the same
 * pass (trainable inducing inputs; GPX_KIND_MERCER_M12 / GPX_KIND_MATERN32 only). */
int gpx_kernel_grad(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB, int divB,
                    const double* hyp, int P, int Q, const double* featA, const double* featB, const double* Kbar,
                    long long strideK, int ldk, double* dhyp, int need_ef, const double* epi_col,
                    const double* epi_rowv, const double* epi_colv, double epi_alpha, double* dptsA, int batch,
                    void* stream);

/* The same hyper-parameter gradient for GPX_KIND_MERCER_M12 when the row points lie ON the column grid: ptsB[w] is a
 * uniform sample grid with spacing delta[w] and ptsA[r][m] == ptsB[r / (divB / divA)][izA[r][m]] exactly (every gpitch
 * caller: windows of a 16 kHz track, inducing points = decimated samples or init_liv maxima, gpitch/init_models.py:9-51).
 * One streaming pass bins Kbar exp(-r) by the integer lag izA[m] - n (the exponential in the caller's distance mode,
 * the cosine mixture as a function of the lag), an O(lags x Q) tail forms the sums: cost independent of Q, no atomics.
 *   izA   [batch / divA, nA] int   grid index of every row point;  delta [batch / divB] grid spacing per window
 *   work  [batch * P * (nB + 2 * nlag)] doubles scratch;  nlag >= nB + max(izA)
 *   dhyp  [batch, P, 2 + 2Q] out, overwritten.  Epilogue arguments as gpx_kernel_grad. */
long long gpx_kernel_grad_lag_workspace_bytes(int nB, int P, int nlag, int batch);   /* size of `work`, -1 on bad arguments */
int gpx_kernel_grad_lag(int mode, const double* ptsA, int nA, int divA, const int* izA, const double* ptsB, int nB, int divB,
                        const double* delta, const double* hyp, int P, int Q, const double* Kbar, long long strideK,
                        int ldk, double* dhyp, int need_ef, const double* epi_col, const double* epi_rowv,
                        const double* epi_colv, double epi_alpha, double* work, int nlag, int batch, void* stream);

/* Inducing points on the window's sample grid (z_j = x[iz_j], izA as in gpx_kernel_grad_lag; iz_j < 0 = pad point of a ragged
 * set): K(z, z) is a column gather of K(z, x), so the second builder launch of gpitch/sgpr_ss.py:42-43 is not needed, and its
 * adjoint is a column scatter, so ONE gradient pass over Kuf_bar covers both matrices.
 *   gpx_kuu_from_kuf:          Kuu[b, m, j] = Kuf[b, m, iz_j] (+ jitter if m == j); pad rows / columns = e_j * pad_diag[b]
 *   gpx_kuu_bar_into_kuf_bar:  Kuf_bar[b, m, iz_j] += Kuu_bar[b, m, j]   (in place; pads receive nothing)
 * Kuf / Kuf_bar [batch, M, ldf] (batch stride strideF), Kuu / Kuu_bar [batch, M, M] contiguous, iz [batch / div, M]. */
int gpx_kuu_from_kuf(const double* Kuf, long long strideF, int ldf, const int* iz, int div, int M, const double* pad_diag,
                     double jitter, double* Kuu, int batch, void* stream);
int gpx_kuu_bar_into_kuf_bar(const double* Kuu_bar, const int* iz, int div, int M, double* Kuf_bar, long long strideF, int ldf,
                             int batch, void* stream);

/* Gradient w.r.t. the row points (inducing inputs)  dptsA[b,m] = sum_p sum_n Kbar[b,m,n] d k_p(z_m, x_n)/d z_m.
 * Replaces tf.gradients w.r.t. Pdgp.za / Pdgp.zc when they are left trainable (gpitch/pdgp.py:80-85 creates them
 * as Params; demos/scripts/demo-modgp.py:40-41 fixes them).  For K(z, z) pass Kbar + Kbar^T (both arguments move).
 *   dptsA [batch, nA] out, overwritten; the caller sums entries that share a point row (divA > 1).
 *   epi_*: the same optional adjoint epilogue as gpx_kernel_grad.
 * Kinds: GPX_KIND_MERCER_M12, GPX_KIND_MATERN32. */
int gpx_kernel_grad_points(int kind, int mode, const double* ptsA, int nA, int divA, const double* ptsB, int nB,
                           int divB, const double* hyp, int P, int Q, const double* featA, const double* featB,
                           const double* Kbar, long long strideK, int ldk, double* dptsA, const double* epi_col,
                           const double* epi_rowv, const double* epi_colv, double epi_alpha, int batch, void* stream);

/* Batched Cholesky + inverse of the factor.  Replaces tf.cholesky (gpitch/sgpr_ss.py:44,51,89; GPflow
 * conditional()) and, through L^-1, every tf.matrix_triangular_solve (gpitch/sgpr_ss.py:48,53,90,94).
 *   A    [batch, M, lda] in: symmetric (lower triangle read); out: L, upper triangle zeroed
 *   Linv [batch, M, ldi] out: L^-1 (lower), upper triangle zeroed
 *   work [batch, 64, M]  scratch;  info [batch] int out */
long long gpx_potrf_workspace_bytes(int M, int batch);                       /* size of `work` below, -1 on bad arguments */
int gpx_potrf_trinv(double* A, long long strideA, int lda, double* Linv, long long strideI, int ldi, double* work,
                    int* info, int M, int batch, void* stream);

/* Generic batched GEMM on the FP64 tensor pipe (DMMA):
 *   C = colscale[n] * (alpha * alpha_vec[b] * op(A) diag(kweight) op(B) + gamma * gamma_vec[b] * Aux[m,n]) + rowvec[m] colvec[n]
 *       + beta * C
 * with triangular k-range skipping.  Replaces tf.matmul / tf.matrix_triangular_solve(L, .) = L^-1 (.) of
 * gpitch/sgpr_ss.py:48-53 and GPflow conditional().  Null pointers disable the optional terms. */
typedef struct {
  const double* A;
  const double* B;
  double* C;
  long long sA, sB, sC;
  int lda, ldb, ldc;
  int M, N, K, batch;
  int flags;
  double alpha, beta, gamma;
  const double* alpha_vec;
  const double* kweight;
  long long sKw;
  const double* Aux;
  long long sAux;
  int ldaux;
  const double* colscale;
  const double* rowvec;
  const double* colvec;
  long long sColscale, sRowvec, sColvec;
  const double* gamma_vec; /* [batch] or NULL: per-batch factor on gamma */
} gpx_gemm_args;
int gpx_gemm(const gpx_gemm_args* args, void* stream);

/* Predictive-marginal epilogue of GPflow conditional(): fmean = A^T q_mu and
 *   mode 0: fvar = Kdiag - sum_m A^2 + sum_m LTA^2   (LTA may be null)
 *   mode 1: fvar = Kdiag + sum_m A o LTA              (G-form: A = Kmn, LTA = G Kmn, q_mu = L^-T q_mu)
 * A, LTA [batch, M, ld] (batch stride strideA), q_mu [batch, M], kdiag [batch], fmean / fvar [batch, N] out. */
int gpx_cond_colstats(const double* A, const double* LTA, long long strideA, int ld, const double* q_mu,
                      const double* kdiag, double* fmean, double* fvar, int M, int N, int batch, int mode,
                      void* stream);

/* out[b,m,n] = alpha * colscale[b,n] * T[b,m,n] + rowvec[b,m] * colvec[b,n]  -- Kbar_mn = 2 T diag(vbar) + a mbar^T of
 * the G-form backward pass (replaces the tf.gradients ops through GPflow conditional()).  T, out [batch, M, ld]. */
int gpx_scale_rank1(const double* T, long long strideT, int ld, const double* colscale, const double* rowvec,
                    const double* colvec, double alpha, double* out, int M, int N, int batch, void* stream);

/* out[b,m] = sum_n A[b,m,n] v[b,n]   (tf.matmul(A, err) of gpitch/sgpr_ss.py:52; mu_bar = A m_bar). */
int gpx_rowdot(const double* A, long long strideA, int ld, const double* v, long long strideV, double* out, int M,
               int N, int batch, void* stream);

/* MpdLik.variational_expectations forward + analytic backward (gpitch/likelihoods.py:33-68,422-447).
 *   Fmu, Fvar, dFmu, dFvar [W, 2P, N] (rows 0..P-1 = activations g_i, rows P..2P-1 = components f_i)
 *   Y [W, N], noise [W];  ve_sum [W] = sum_n var_exp, dnoise [W]  (both overwritten)
 *   ve_pointwise [W, N] optional (null to skip); dFmu/dFvar/dnoise null to skip the backward pass. */
int gpx_varexp(const double* Fmu, const double* Fvar, const double* Y, const double* noise, int P, int W, int N,
               int nlin, double* ve_sum, double* dFmu, double* dFvar, double* dnoise, double* ve_pointwise,
               void* stream);

/* Hann overlap-add of per-window predictions on the device: merged_mean (win = scipy hann(ws)) and merged_variance
 * (win = hann(ws)^2) of gpitch/window_overlap.py:19-59, same index arithmetic and rounding (bit-exact).
 *   Y [num_windows, ws] windows at hop (ws-1)/2;  win [ws] (device);  out [n] */
int gpx_overlap_add(const double* Y, const double* win, int num_windows, int ws, int n, double* out, void* stream);

/* gpflow.kullback_leiblers.gauss_kl(q_mu, q_sqrt) with K=None (whitened; gpitch/pdgp.py:120-121) and its
 * gradient.  q_mu [batch, M], q_sqrt [batch, M, M] (lower triangle used), kl [batch], dmu [batch, M],
 * dLq [batch, M, M] (upper triangle zero).  dmu / dLq may be null. */
int gpx_gauss_kl_white(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* dmu,
                       double* dLq, void* stream);

/* Composite entry point: SGPRSS.build_likelihood (gpitch/sgpr_ss.py:29-71) and its gradient for W windows in ONE call --
 * what GPflow's Model._objective evaluates for this model (value and d/d(constrained parameters); the host applies the
 * free-state chain rule) -- for hosts without torch.  The same launch sequence as gpitch_b200/functions.py:SGPRBound.
 *   x, y [W, N]; z [W, M]; hyp [W, P, 2+2Q]; noise [W]; reg != 0 adds -1000 sum_p |variance_p| (sgpr_ss.py:64-68)
 *   iz [W, M], delta [W], nlag: optional (NULL / 0) grid structure of the inducing points as in gpx_kernel_grad_lag ->
 *       K(z, z) gathered from K(z, x) and one lag-histogram gradient pass over Kuf_bar + scattered Kuu_bar instead of a second
 *       builder launch and two per-element gradient passes
 *   bound [W] out; dhyp [W, P, 2+2Q], dnoise [W] out (dhyp NULL = value only; need_ef = 0 leaves the energy / frequency
 *       columns of dhyp zero: fixed partials)
 *   info [2 * W] int out: LAPACK-style status of chol(Kuu) (first W) and chol(I + A A^T) (next W)
 *   work: gpx_sgpr_bound_workspace_bytes(...) bytes of device scratch (with_grad = dhyp != NULL; nlag as passed) */
long long gpx_sgpr_bound_workspace_bytes(int kind, int N, int M, int P, int Q, int W, int with_grad, int nlag);
int gpx_sgpr_bound(int kind, int mode, const double* x, const double* y, const double* z, int N, int M, int W,
                   const double* hyp, int P, int Q, const double* noise, double jitter, int reg, int need_ef, const int* iz,
                   const double* delta, int nlag, double* bound, double* dhyp, double* dnoise, int* info, double* work,
                   void* stream);

/* Packed lower triangles for the host-facing path: packed [batch, M (M + 1) / 2] (row-major: (i, j <= i) at
 * i (i + 1) / 2 + j) <-> dense [batch, M, M] (unpack writes exact zeros above the diagonal).  The reference stores q_sqrt
 * as a dense M x M Param but reads only tf.matrix_band_part(q_sqrt, -1, 0) (GPflow conditional / gauss_kl via
 * gpitch/pdgp.py:120-155): the strict upper triangle and its gradient never carry information. */
int gpx_tril_unpack(const double* packed, double* dense, int M, int batch, void* stream);
int gpx_tril_pack(const double* dense, double* packed, int M, int batch, void* stream);

/* The same KL, returning next to it the masked factor tril_out = tf.matrix_band_part(q_sqrt, -1, 0) [batch, M, M] that
 * GPflow conditional() multiplies with (one pass over q_sqrt instead of a separate masking copy).  The gradient is
 * analytic: d kl / d q_mu = q_mu, d kl / d q_sqrt = tril_out - diag(1 / diag(tril_out)). */
int gpx_gauss_kl_white_tril(const double* q_mu, const double* q_sqrt, int M, int batch, double* kl, double* tril_out,
                            void* stream);

/* Measurement helpers (bench.py): number of kernels this library has launched so far in the process, and the
 * FP64 tensor-pipe peak of the current device (register-resident mma.sync m8n8k4 loop, best of reps, TFLOP/s
 * written to the HOST double *tflops; synchronises). */
unsigned long long gpx_launch_count(void);
/* gpx_gemm launches that took the TMA + mbarrier kernel (csrc/gemm_tma.cu) rather than the cp.async kernel kept for
 * operands that miss TMA's 16-byte alignment rules (odd leading dimension / unaligned base). */
unsigned long long gpx_gemm_tma_launch_count(void);
int gpx_dmma_peak(int reps, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif

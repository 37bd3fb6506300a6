"""[GPflow-0.5, recalled] semantics used by the gpitch hot path, restated in torch fp64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  GPflow 0.5 / TF 1.2.1 are not on disk; every
function here follows SURVEY.md Appendix A (A.1-A.7) and cites the reference call site that relies
on it.  PARITY UNPINNED for this file (no reference golden vectors exist for it).
"""
import numpy as np
import torch

JITTER = 1e-6          # gpflow.settings.numerics.jitter_level (recalled)
DIST_EPS = 1e-12       # Stationary.euclid_dist adds 1e-12 under the sqrt (recalled)
POSITIVE_LOWER = 1e-6  # transforms.positive == Log1pe(lower=1e-6) (recalled)
DTYPE = torch.float64  # settings.dtypes.float_type


def as_t(x):
    if isinstance(x, torch.Tensor):
        return x.to(DTYPE)
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


# ---------------------------------------------------------------- A.1 Stationary
def square_dist(X, X2, lengthscales):
    """Stationary.square_dist: distance by expansion, evaluation order ((-2*P)+Xs_i)+X2s_j.
    Call sites: matern12_spectral_mixture.py:106 (through euclid_dist), init_kernels.py:12."""
    X = X / lengthscales
    Xs = torch.sum(torch.square(X), 1)
    if X2 is None:
        return -2 * torch.matmul(X, X.t()) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
    X2 = X2 / lengthscales
    X2s = torch.sum(torch.square(X2), 1)
    return -2 * torch.matmul(X, X2.t()) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)


CLEAN_LENGTHSCALE_GRAD = False


class _EuclidDistCleanGrad(torch.autograd.Function):
    """Same forward VALUE as euclid_dist (reference operation order), but d r / d lengthscale evaluated from the
    closed form -s / (l r) instead of reverse-mode through the expansion.  Reverse-mode through
    (x/l)^2 - 2 (x/l)(x'/l) + (x'/l)^2 sums three O(x~^2 / l) terms that cancel to O(d~^2 / l); in fp64 that loses
    ~log10(x~^2 / d~^2) digits and the error depends on the GEMV summation order, i.e. the reference's own
    lengthscale gradient is not reproducible beyond ~1e-5 at t = 10 s (tests/test_formulas_cpu.py::
    test_lengthscale_grad_noise_of_reference checks this against 50-digit arithmetic).  Tests that need a 1e-8
    target for the lengthscale gradient at absolute time stamps switch this on."""

    @staticmethod
    def forward(ctx, X, X2, lengthscales):
        s = square_dist(X, X2, lengthscales)
        r = torch.sqrt(s + DIST_EPS)
        ctx.save_for_backward(s, r, lengthscales)
        return r

    @staticmethod
    def backward(ctx, g):
        s, r, l = ctx.saved_tensors
        return None, None, (g * (-s / (l * r))).sum().reshape(l.shape)


def euclid_dist(X, X2, lengthscales):
    """Stationary.euclid_dist = sqrt(square_dist + 1e-12)."""
    if CLEAN_LENGTHSCALE_GRAD and isinstance(lengthscales, torch.Tensor) and lengthscales.requires_grad:
        return _EuclidDistCleanGrad.apply(X, X2, lengthscales)
    return torch.sqrt(square_dist(X, X2, lengthscales) + DIST_EPS)


def matern32_K(X, X2, variance, lengthscales):
    """gpflow.kernels.Matern32.K (activation kernel, init_kernels.py:12, demo-modgp.py:32)."""
    r = euclid_dist(X, X2, lengthscales)
    return variance * (1. + np.sqrt(3.) * r) * torch.exp(-np.sqrt(3.) * r)


def stationary_Kdiag(X, variance):
    """Stationary.Kdiag = fill([N], variance)."""
    return torch.ones(X.shape[0], dtype=DTYPE) * torch.squeeze(variance)


# ---------------------------------------------------------------- A.3 transforms
def positive_forward(x):
    """transforms.positive: y = softplus(x) + 1e-6."""
    return torch.nn.functional.softplus(x) + POSITIVE_LOWER


def positive_backward(y):
    """x = log(exp(y - 1e-6) - 1)."""
    y = as_t(y) - POSITIVE_LOWER
    return y + torch.log(-torch.expm1(-y))


def logistic_forward(x, a=0., b=1.):
    return a + (b - a) * torch.sigmoid(x)


# ---------------------------------------------------------------- A.5 conditional
def conditional(Xnew, X, K_fn, Kdiag_fn, f, q_sqrt=None, whiten=False, jitter=JITTER):
    """gpflow.conditionals.conditional(Xnew, X, kern, f, full_cov=False, q_sqrt, whiten).
    Call sites: pdgp.py:147-155, :176, :185, :199-203.
    K_fn(A, B_or_None) -> kernel matrix; Kdiag_fn(A) -> [N].
    f: [M, K]; q_sqrt: [M, M, K] (lower triangle taken) or [M, K] or None.
    Returns fmean [N, K], fvar [N, K]."""
    num_data = X.shape[0]
    num_func = f.shape[1]
    Kmn = K_fn(X, Xnew)
    Kmm = K_fn(X, None) + torch.eye(num_data, dtype=DTYPE) * jitter
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = Kdiag_fn(Xnew) - torch.sum(torch.square(A), 0)
    fvar = fvar[None, :].repeat(num_func, 1)                      # K x N
    if not whiten:
        A = torch.linalg.solve_triangular(Lm.t(), A, upper=True)
    fmean = torch.matmul(A.t(), f)
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A[None, :, :] * q_sqrt.t()[:, :, None]           # K x M x N
        else:
            L = torch.tril(q_sqrt.permute(2, 0, 1))                # K x M x M
            A_tiled = A[None, :, :].repeat(num_func, 1, 1)
            LTA = torch.matmul(L.transpose(1, 2), A_tiled)         # K x M x N
        fvar = fvar + torch.sum(torch.square(LTA), 1)
    return fmean, fvar.t()


# ---------------------------------------------------------------- A.6 gauss_kl
def gauss_kl(q_mu, q_sqrt, K=None):
    """gpflow.kullback_leiblers.gauss_kl(q_mu [M,L], q_sqrt [M,M,L], K=None).
    Call sites: pdgp.py:120-129."""
    M, num_latent = q_mu.shape
    if K is None:
        alpha = q_mu
    else:
        Lp = torch.linalg.cholesky(K)
        alpha = torch.linalg.solve_triangular(Lp, q_mu, upper=False)
    Lq = torch.tril(q_sqrt.permute(2, 0, 1))                       # L x M x M
    twoKL = torch.sum(torch.square(alpha))                         # Mahalanobis
    twoKL = twoKL - float(M * num_latent)                          # constant
    twoKL = twoKL - torch.sum(torch.log(torch.square(torch.diagonal(Lq, dim1=1, dim2=2))))
    if K is None:
        twoKL = twoKL + torch.sum(torch.square(Lq))                # trace
    else:
        LpiLq = torch.linalg.solve_triangular(Lp[None].expand(num_latent, M, M), Lq, upper=False)
        twoKL = twoKL + torch.sum(torch.square(LpiLq))
        twoKL = twoKL + num_latent * torch.sum(torch.log(torch.square(torch.diagonal(Lp))))
    return 0.5 * twoKL


# ---------------------------------------------------------------- A.7 quadrature
def hermgauss(n):
    """gpflow.quadrature.hermgauss == np.polynomial.hermite.hermgauss cast to float64."""
    x, w = np.polynomial.hermite.hermgauss(n)
    return x.astype(np.float64), w.astype(np.float64)


# ---------------------------------------------------------------- A.4 SGPR.build_predict
def sgpr_build_predict(X, Y, Z, Xnew, K_fn, Kdiag_fn, noise_var, full_cov=False, jitter=JITTER):
    """gpflow.sgpr.SGPR.build_predict (inherited predict_f of SGPRSS; separation.py:306).
    Recomputes err, Kuf, Kuu, L, A, B, LB, c exactly as sgpr_ss.py:40-53."""
    num_inducing = Z.shape[0]
    err = Y
    Kuf = K_fn(Z, X)
    Kuu = K_fn(Z, None) + torch.eye(num_inducing, dtype=DTYPE) * jitter
    Kus = K_fn(Z, Xnew)
    sigma = torch.sqrt(noise_var)
    L = torch.linalg.cholesky(Kuu)
    A = torch.linalg.solve_triangular(L, Kuf, upper=False) / sigma
    B = torch.matmul(A, A.t()) + torch.eye(num_inducing, dtype=DTYPE)
    LB = torch.linalg.cholesky(B)
    Aerr = torch.matmul(A, err)
    c = torch.linalg.solve_triangular(LB, Aerr, upper=False) / sigma
    tmp1 = torch.linalg.solve_triangular(L, Kus, upper=False)
    tmp2 = torch.linalg.solve_triangular(LB, tmp1, upper=False)
    mean = torch.matmul(tmp2.t(), c)
    if full_cov:
        var = K_fn(Xnew, None) + torch.matmul(tmp2.t(), tmp2) - torch.matmul(tmp1.t(), tmp1)
        var = var[:, :, None].repeat(1, 1, Y.shape[1])
    else:
        var = Kdiag_fn(Xnew) + torch.sum(torch.square(tmp2), 0) - torch.sum(torch.square(tmp1), 0)
        var = var.reshape(-1, 1).repeat(1, Y.shape[1])
    return mean, var

"""torch.autograd.Function wrappers over the C-ABI kernels (gpitch_b200/_lib.py).

Each Function is one stage of the variational-GP inner loop, batched over (window, latent GP), with a
hand-derived analytic backward (formulas pinned on CPU by tests/test_formulas_cpu.py) instead
of the generic reverse-mode graph `tf.gradients` builds for the reference (GPflow Model._objective)."""
import math
import torch

from . import _lib as L

LOG2PI = math.log(2.0 * math.pi)


def _matvec(A, v):
    """A v per batch entry (A [b,M,K], v [b,K]) on the row-dot kernel instead of a 1-column GEMM launch."""
    return L.rowdot(A.contiguous(), v.contiguous())


def _matTvec(A, v):
    """A^T v per batch entry (A [b,K,M], v [b,K]) on the column-statistics kernel (its fmean output)."""
    return L.cond_colstats(A.contiguous(), None, v.contiguous(), v.new_zeros(A.shape[0]))[0]


def _eye_add_(A, v):
    A.diagonal(dim1=-2, dim2=-1).add_(v)
    return A


class KernelMatrix(torch.autograd.Function):
    """K[b] = sum_p k_p(ptsA[b // divA], ptsB[b // divB]) (+ jitter I).  hyp [batch, P, 2 + 2Q].
    Replaces Kern.K of gpitch/matern12_spectral_mixture.py:38-56,102-117, GPflow Matern32 and Add."""

    @staticmethod
    def forward(ctx, hyp, ptsA, ptsB, kind, mode, jitter, need_ef, lag=None):
        ctx.lag = lag              # batched.grid_lags(ptsB, ptsA): row points on the column grid -> lag-histogram gradient
        hyp = hyp.contiguous()
        batch, P, HS = hyp.shape
        Q = (HS - 2) // 2
        featA = featB = None
        if kind == 'mercer_m12':
            featA = L.features(ptsA, hyp, P, Q)
            featB = featA if ptsB is ptsA else L.features(ptsB, hyp, P, Q)
        K = L.kernel_build(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, jitter=jitter)
        ctx.save_for_backward(hyp, ptsA, ptsB, featA, featB)
        ctx.cfg = (kind, mode, P, Q, need_ef, ptsB is ptsA)
        return K

    @staticmethod
    def backward(ctx, Kbar):
        hyp, ptsA, ptsB, featA, featB = ctx.saved_tensors
        kind, mode, P, Q, need_ef, same = ctx.cfg
        if Kbar.stride(-1) != 1 or Kbar.stride(-2) < Kbar.shape[-1]:
            Kbar = Kbar.contiguous()
        if ctx.lag is not None and kind == 'mercer_m12' and not ctx.needs_input_grad[1]:
            dhyp = L.kernel_grad_lag(mode, ptsA, ptsB, hyp, P, Q, Kbar, ctx.lag, need_ef=need_ef)
        else:
            dhyp = L.kernel_grad(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, Kbar, need_ef=need_ef)
        dA = None
        if ctx.needs_input_grad[1]:
            # trainable inducing inputs (gpitch/pdgp.py:80-85).  For K(z, z) both arguments move: by symmetry of k the
            # total derivative is the row-point derivative under Kbar + Kbar^T (returned for the first slot only).
            Kz = Kbar + Kbar.transpose(-1, -2) if same else Kbar
            dA = L.kernel_grad_points(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, Kz.contiguous())
            dA = dA.view(ptsA.shape[0], -1, ptsA.shape[1]).sum(1)
        if ctx.needs_input_grad[2] and not same:
            raise NotImplementedError('gradient w.r.t. the column points (data) is not on the gpitch path')
        return dhyp, dA, None, None, None, None, None, None


class KernelPairOnGrid(torch.autograd.Function):
    """(Kuf, Kuu + jitter I) of gpitch/sgpr_ss.py:42-43 when the inducing points lie on the window's sample grid
    (z_j = x[iz_j]; lag = batched.grid_lags(x, z)): K(z, z) is gathered from the columns iz of K(z, x) instead of being
    built by a second launch -- the same kernel function of the same fp64 inputs -- and in the backward pass Kuu_bar is
    scattered into Kuf_bar, after which ONE lag-histogram pass yields the hyper-gradient of both matrices.  Pad points of
    ragged inducing sets (iz < 0) get a decoupled diagonal entry; the collapsed bound does not depend on it."""

    @staticmethod
    def forward(ctx, hyp, z, x, kind, mode, jitter, need_ef, lag):
        assert kind == 'mercer_m12'
        hyp = hyp.contiguous()
        batch, P, HS = hyp.shape
        Q = (HS - 2) // 2
        fz, fx = L.features(z, hyp, P, Q), L.features(x, hyp, P, Q)
        Kuf = L.kernel_build(kind, mode, z, x, hyp, P, Q, fz, fx, jitter=0.0)
        pad_diag = (hyp[:, :, 0] * hyp[:, :, 2:2 + Q].sum(-1)).sum(1)
        Kuu = L.kuu_from_kuf(Kuf, lag[0], pad_diag, jitter)
        ctx.save_for_backward(hyp, z, x)
        ctx.cfg = (mode, P, Q, need_ef)
        ctx.lag = lag
        return Kuf, Kuu

    @staticmethod
    def backward(ctx, dKuf, dKuu):
        hyp, z, x = ctx.saved_tensors
        mode, P, Q, need_ef = ctx.cfg
        Kb = dKuf.contiguous().clone()                       # (the incoming gradient buffer is not ours to modify)
        L.kuu_bar_into_kuf_bar(dKuu.contiguous(), ctx.lag[0], Kb)
        dhyp = L.kernel_grad_lag(mode, z, x, hyp, P, Q, Kb, ctx.lag, need_ef=need_ef)
        return dhyp, None, None, None, None, None, None, None


class SVGPConditional(torch.autograd.Function):
    """GPflow-0.5 conditional(Xnew, X, kern, f, full_cov=False, q_sqrt, whiten=True) (call sites
    gpitch/pdgp.py:147-155,176,185,199-203), from prebuilt Kmn [b,M,N], Kmm (+jitter) [b,M,M], kdiag [b],
    q_mu [b,M], q_sqrt [b,M,M] -> fmean [b,N], fvar [b,N], info [b]."""

    @staticmethod
    def forward(ctx, Kmn, Kmm, kdiag, q_mu, q_sqrt, Lm=None, Linv=None, info=None):
        Lq = torch.tril(q_sqrt)
        Lm, Linv, info = _factor(Kmm, (Lm, Linv, info))
        A = L.gemm(Linv, Kmn, flags=L.GEMM_A_LOWER)
        LTA = L.gemm(Lq, A, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)
        q_mu = q_mu.contiguous()
        fmean, fvar = L.cond_colstats(A, LTA, q_mu, kdiag.contiguous())
        del LTA                                   # only its column norms are needed; the backward pass works from A
        ctx.save_for_backward(Lm, Linv, A, Lq, q_mu)
        ctx.mark_non_differentiable(info)
        return fmean, fvar, info

    @staticmethod
    def backward(ctx, mbar, vbar, _info):
        Lm, Linv, A, Lq, q_mu = ctx.saved_tensors
        mbar, vbar = mbar.contiguous(), vbar.contiguous()
        mubar = L.rowdot(A, mbar)
        SD = L.gemm(A, A, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=vbar)
        W1 = L.gemm(Lq, Lq, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        _eye_add_(W1, -1.0)                                                                   # Lq Lq^T - I (symmetric)
        # Kbar_mn = L^-T Abar,  Abar = mu mbar^T + 2 (Lq Lq^T - I) A diag(vbar)
        #         = 2 [L^-T (Lq Lq^T - I)] A diag(vbar) + (L^-T mu) mbar^T :
        # ONE dense M x M x N product with a fused column-scale / rank-1 epilogue instead of two triangular ones.
        H = L.gemm(Linv, W1, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)
        alpha_vec = _matTvec(Linv, q_mu)
        dKmn = L.gemm(H, A, alpha=2.0, colscale=vbar, rowvec=alpha_vec, colvec=mbar)
        dLq, dKmm = _conditional_tail(Lm, Linv, H, SD, Lq, alpha_vec, mubar)
        return dKmn, dKmm, vbar.sum(1), mubar, dLq, None, None, None


def _factor(Kmm, pre):
    """Cholesky factor, its inverse and the LAPACK-style status of Kmm -- or the ones Unwhiten already computed."""
    if pre is not None and pre[0] is not None:
        return pre
    return L.potrf_trinv(Kmm.clone())


def _build_kmn(hyp, z, x, kind, mode):
    """Kmn = sum_p k_p(z, x) for the fused conditional() stages, which own it: its adjoint 2 T diag(vbar) + a mbar^T is
    never written to memory but consumed from T by the builder-gradient kernels' fused epilogue."""
    hyp = hyp.contiguous()
    P, Q = hyp.shape[1], (hyp.shape[2] - 2) // 2
    fz = fx = None
    if kind == 'mercer_m12':
        fz, fx = L.features(z, hyp, P, Q), L.features(x, hyp, P, Q)
    return hyp, L.kernel_build(kind, mode, z, x, hyp, P, Q, fz, fx, jitter=0.0), fz, fx, P, Q


def _kmn_backward(ctx, hyp, z, x, fz, fx, T, epilogue):
    """Hyper-parameter (and, if z is trainable, inducing-input) gradients of the Kmn a conditional() stage built."""
    kind, mode, P, Q, need_ef = ctx.cfg
    if not ctx.needs_input_grad[1]:
        if getattr(ctx, 'lag', None) is not None and kind == 'mercer_m12':     # inducing points on the sample grid
            return L.kernel_grad_lag(mode, z, x, hyp, P, Q, T, ctx.lag, need_ef=need_ef, epilogue=epilogue), None
        return L.kernel_grad(kind, mode, z, x, hyp, P, Q, fz, fx, T, need_ef=need_ef, epilogue=epilogue), None
    dhyp, dz = L.kernel_grad(kind, mode, z, x, hyp, P, Q, fz, fx, T, need_ef=need_ef, epilogue=epilogue, with_points=True)
    return dhyp, dz.view(z.shape[0], -1, z.shape[1]).sum(1)


def _conditional_tail(Lm, Linv, H, SD, Lq, alpha_vec, mubar, klbar=None):
    """Shared M x M part of the conditional() backward passes: dLq and Kmm_bar (Cholesky adjoint) from S_D = A D A^T.
    klbar [b]: the stage also owns the whitened KL term of this latent GP -- its gradient klbar (Lq - diag(1 / diag Lq))
    rides in the epilogue of the dLq product (Aux = Lq) instead of being formed, scaled and accumulated by three
    element-wise passes over [b, M, M]."""
    if klbar is None:
        dLq = L.gemm(SD, Lq, flags=L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER, alpha=2.0)
    else:
        dLq = L.gemm(SD, Lq, flags=L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER, alpha=2.0, aux=Lq, gamma=1.0,
                     gamma_vec=klbar)
        dLq.diagonal(dim1=1, dim2=2).sub_(klbar[:, None] / Lq.diagonal(dim1=1, dim2=2))
    # Lbar = -tril(L^-T Abar A^T),  Abar A^T = mu mubar^T + 2 (Lq Lq^T - I) S_D  =>  L^-T Abar A^T = 2 H S_D + alpha mubar^T
    Lbar = L.gemm(H, SD, flags=L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER, alpha=-2.0, rowvec=(-alpha_vec).contiguous(), colvec=mubar)
    # P = Phi(L^T Lbar) (lower triangle, halved diagonal); P + P^T is the lower triangle of L^T Lbar mirrored, with the
    # diagonal counted once -- exactly the C_LOWER | C_MIRROR output of the product
    Psym = L.gemm(Lm, Lbar, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
    U = L.gemm(Psym, Linv, flags=L.GEMM_B_LOWER)
    dKmm = L.gemm(Linv, U, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, alpha=0.5)   # symmetric
    return dLq, dKmm


class SVGPConditionalHA(torch.autograd.Function):
    """Same function as SVGPConditional with THREE M^2 N-class products instead of four, and the same backward
    stability (no explicit Kmm^-1-like matrix is ever formed): with A = L^-1 Kmn, H = L^-T (Lq Lq^T - I), a = L^-T q_mu,
        T = H A (= G Kmn),   fvar = Kdiag + sum_m Kmn o T,   fmean = Kmn^T a,
        Kbar_mn = 2 T diag(vbar) + a mbar^T (fused into the builder-gradient kernels, never written),
        S_D = A diag(vbar) A^T (weighted SYRK on A itself).
    These two forms own their Kmn and the whitened KL term of their latent GPs (gauss_kl(q_mu, q_sqrt), pdgp.py:120-121):
    inputs are (hyp, z, x, Kmm, kdiag, q_mu, q_sqrt, kind, mode, need_ef), outputs (fmean, fvar, kl, info).
    Forward: one triangular + one dense product; backward: one SYRK.  Measured against an 80-bit evaluation on a
    jitter-dominated Matern-3/2 group (cond(Kmm) = 7e8): fvar error 1.7e-11 (triangular form 1.7e-11, G-form 8e-8)."""

    @staticmethod
    def forward(ctx, hyp, z, x, Kmm, kdiag, q_mu, q_sqrt, kind, mode, need_ef, Lm=None, Linv=None, info=None, lag=None):
        ctx.lag = lag
        q_mu = q_mu.contiguous()
        kl, Lq = L.gauss_kl_white_tril(q_mu, q_sqrt.contiguous())       # gauss_kl(q_mu, q_sqrt) + tril(q_sqrt) in one pass
        Lm, Linv, info = _factor(Kmm, (Lm, Linv, info))
        W1 = L.gemm(Lq, Lq, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        _eye_add_(W1, -1.0)
        H = L.gemm(Linv, W1, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)                       # L^-T (Lq Lq^T - I)
        alpha_vec = _matTvec(Linv, q_mu)
        hyp, Kmn, fz, fx, P, Q = _build_kmn(hyp, z, x, kind, mode)
        A = L.gemm(Linv, Kmn, flags=L.GEMM_A_LOWER)
        T = L.gemm(H, A)
        fmean, fvar = L.cond_colstats(Kmn, T, alpha_vec, kdiag.contiguous(), mode=1)
        ctx.save_for_backward(Lm, Linv, A, T, Lq, H, alpha_vec, hyp, z, x, fz, fx, q_mu)
        ctx.cfg = (kind, mode, P, Q, need_ef)
        ctx.mark_non_differentiable(info)
        return fmean, fvar, kl, info

    @staticmethod
    def backward(ctx, mbar, vbar, klbar, _info):
        Lm, Linv, A, T, Lq, H, alpha_vec, hyp, z, x, fz, fx, q_mu = ctx.saved_tensors
        mbar, vbar, klbar = mbar.contiguous(), vbar.contiguous(), klbar.contiguous()
        dhyp, dz = _kmn_backward(ctx, hyp, z, x, fz, fx, T, (2.0, vbar, alpha_vec, mbar))
        mubar = L.rowdot(A, mbar)
        SD = L.gemm(A, A, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=vbar)
        dLq, dKmm = _conditional_tail(Lm, Linv, H, SD, Lq, alpha_vec, mubar, klbar)
        return dhyp, dz, None, dKmm, vbar.sum(1), mubar + klbar[:, None] * q_mu, dLq, None, None, None, None, None, None, None


class SVGPConditionalG(torch.autograd.Function):
    """Same function as SVGPConditional, "G-form": with G = L^-T (Lq Lq^T - I) L^-1 and a = L^-T q_mu,
        fvar = Kdiag + sum_m Kmn o (G Kmn),   fmean = Kmn^T a,
    so the forward pass is ONE dense M x M x N product (T = G Kmn) and the backward pass reuses it:
        Kbar_mn = 2 T diag(vbar) + a mbar^T  (never materialised: the builder-gradient kernels apply it to T on the fly),
        Gbar = Kmn diag(vbar) Kmn^T  (one weighted SYRK).
    2 M^2 N-class products instead of 4.  It forms Kmm^-1-like matrices explicitly, so its rounding error grows like
    eps * cond(Kmm) (measured 6e-17 * cond on the C3 kernels); BatchedPdgp only selects it for groups whose Cholesky
    factors certify cond(Kmm) <~ 1e4 (every MercerMatern12sm component group of the named configs), keeping the
    result within the 1e-8 parity budget.  Jitter-dominated groups (Matern32 activations, cond ~ 1e9) stay on
    SVGPConditional."""

    @staticmethod
    def forward(ctx, hyp, z, x, Kmm, kdiag, q_mu, q_sqrt, kind, mode, need_ef, Lm=None, Linv=None, info=None, lag=None):
        ctx.lag = lag
        q_mu = q_mu.contiguous()
        kl, Lq = L.gauss_kl_white_tril(q_mu, q_sqrt.contiguous())       # gauss_kl(q_mu, q_sqrt) + tril(q_sqrt) in one pass
        Lm, Linv, info = _factor(Kmm, (Lm, Linv, info))
        W1 = L.gemm(Lq, Lq, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        _eye_add_(W1, -1.0)
        H = L.gemm(Linv, W1, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)                       # L^-T (Lq Lq^T - I)
        G = L.gemm(H, Linv, flags=L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)       # symmetric
        alpha_vec = _matTvec(Linv, q_mu)
        hyp, Kmn, fz, fx, P, Q = _build_kmn(hyp, z, x, kind, mode)
        T = L.gemm(G, Kmn)
        fmean, fvar = L.cond_colstats(Kmn, T, alpha_vec, kdiag.contiguous(), mode=1)
        ctx.save_for_backward(Lm, Linv, Kmn, T, Lq, H, alpha_vec, hyp, z, x, fz, fx, q_mu)
        ctx.cfg = (kind, mode, P, Q, need_ef)
        ctx.mark_non_differentiable(info)
        return fmean, fvar, kl, info

    @staticmethod
    def backward(ctx, mbar, vbar, klbar, _info):
        Lm, Linv, Kmn, T, Lq, H, alpha_vec, hyp, z, x, fz, fx, q_mu = ctx.saved_tensors
        mbar, vbar, klbar = mbar.contiguous(), vbar.contiguous(), klbar.contiguous()
        dhyp, dz = _kmn_backward(ctx, hyp, z, x, fz, fx, T, (2.0, vbar, alpha_vec, mbar))
        abar = L.rowdot(Kmn, mbar)                                                          # d / d alpha = Kmn mbar
        mubar = _matvec(Linv, abar)   # = A mbar
        Gbar = L.gemm(Kmn, Kmn, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=vbar)
        U1 = L.gemm(Linv, Gbar, flags=L.GEMM_A_LOWER)
        SD = L.gemm(U1, Linv, flags=L.GEMM_TRANS_B | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)   # A D A^T
        dLq, dKmm = _conditional_tail(Lm, Linv, H, SD, Lq, alpha_vec, mubar, klbar)
        return dhyp, dz, None, dKmm, vbar.sum(1), mubar + klbar[:, None] * q_mu, dLq, None, None, None, None, None, None, None


class Unwhiten(torch.autograd.Function):
    """Non-whitened variational parameters -> their whitened equivalents (GPflow conditional / gauss_kl with
    whiten=False, gpitch/pdgp.py:122-129): q(u) = N(q_mu, Lq Lq^T) on u equals the whitened model at
        mu_w = L^-1 q_mu,   Lq_w = L^-1 Lq   (lower x lower = lower),   L = chol(Kmm + jitter I),
    for the conditional (A <- L^-T A) as well as for the KL (alpha = L^-1 q_mu, trace ||L^-1 Lq||^2, and
    log diag(Lq_w)^2 = log diag(Lq)^2 - log diag(L)^2).  Only M x M work."""

    @staticmethod
    def forward(ctx, q_mu, q_sqrt, Kmm):
        Lq = torch.tril(q_sqrt)
        Lm, Linv, info = L.potrf_trinv(Kmm.clone())
        mu_w = _matvec(Linv, q_mu)
        Lq_w = L.gemm(Linv, Lq, flags=L.GEMM_A_LOWER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER)
        ctx.save_for_backward(Lm, Linv, q_mu, Lq)
        ctx.mark_non_differentiable(Lm, Linv, info)        # handed to conditional(): one factorisation per Kmm
        return mu_w, Lq_w, Lm, Linv, info

    @staticmethod
    def backward(ctx, mubar_w, Lqbar_w, _Lm, _Linv, _info):
        Lm, Linv, q_mu, Lq = ctx.saved_tensors
        Lqbar_w = torch.tril(Lqbar_w).contiguous()
        mubar_w = mubar_w.contiguous()
        dq_mu = _matTvec(Linv, mubar_w)
        dLq = L.gemm(Linv, Lqbar_w, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER)
        # Linv_bar = mubar_w q_mu^T + Lqbar_w Lq^T;   L_bar = -tril(L^-T Linv_bar L^-T);   Kmm_bar by the Cholesky adjoint
        Lib = L.gemm(Lqbar_w, Lq, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER, rowvec=mubar_w, colvec=q_mu.contiguous())
        T1 = L.gemm(Linv, Lib, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)
        Lbar = L.gemm(T1, Linv, flags=L.GEMM_TRANS_B | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER, alpha=-1.0)
        Psym = L.gemm(Lm, Lbar, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)   # Phi(.) + Phi(.)^T
        U = L.gemm(Psym, Linv, flags=L.GEMM_B_LOWER)
        dKmm = L.gemm(Linv, U, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, alpha=0.5)
        return dq_mu, dLq, dKmm


def cholesky_cond_estimate(Kmm):
    """Lower bound of cond_2(Kmm) per batch entry from its Cholesky factor: (max_i L_ii / min_i L_ii)^2."""
    Lm, _, info = L.potrf_trinv(Kmm.clone())
    d = Lm.diagonal(dim1=1, dim2=2)
    est = (d.max(1).values / d.min(1).values) ** 2
    return torch.where(info == 0, est, torch.full_like(est, float('inf')))


class SGPRBound(torch.autograd.Function):
    """Collapsed (Titsias) bound of SGPRSS.build_likelihood, gpitch/sgpr_ss.py:40-62, batched over windows.
    Kuf [b,M,N], Kuu (+jitter) [b,M,M], sum_kdiag [b], y [b,N], noise [b] -> bound [b], info [b,2]."""

    @staticmethod
    def forward(ctx, Kuf, Kuu, sum_kdiag, y, noise):
        b, M, N = Kuf.shape
        y = y.contiguous()
        inv_sigma = torch.rsqrt(noise).contiguous()
        Lm, Linv, info1 = L.potrf_trinv(Kuu.clone())
        A = L.gemm(Linv, Kuf, flags=L.GEMM_A_LOWER, alpha_vec=inv_sigma)
        B = L.gemm(A, A, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        trAAT = B.diagonal(dim1=1, dim2=2).sum(1)
        _eye_add_(B, 1.0)
        LB, LBinv, info2 = L.potrf_trinv(B.clone())
        Aerr = L.rowdot(A, y)
        c = _matvec(LBinv, Aerr) * inv_sigma[:, None]
        yy = (y * y).sum(1)
        bound = (-0.5 * N * LOG2PI - torch.log(LB.diagonal(dim1=1, dim2=2)).sum(1) - 0.5 * N * torch.log(noise)
                 - 0.5 * yy / noise + 0.5 * (c * c).sum(1) - 0.5 * sum_kdiag / noise + 0.5 * trAAT)
        info = torch.stack([info1, info2], 1)
        ctx.save_for_backward(Linv, A, B, LBinv, c, y, noise, inv_sigma, sum_kdiag, yy)
        ctx.mark_non_differentiable(info)
        return bound, info

    @staticmethod
    def backward(ctx, gbar, _info):
        Linv, A, B, LBinv, c, y, noise, inv_sigma, sum_kdiag, yy = ctx.saved_tensors
        b, M, N = A.shape
        v1 = _matTvec(LBinv, c).contiguous()                                                  # B^-1 A u
        Binv = L.gemm(LBinv, LBinv, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        Atv = _matTvec(A, v1)
        u = y * inv_sigma[:, None]
        w = (u - Atv).contiguous()
        # Kbar_uf = L^-T Abar / sigma,  Abar = (I - B^-1) A + v w^T
        #         = [L^-T (I - B^-1)] A / sigma + (L^-T v) w^T / sigma : one dense product with a rank-1 epilogue.
        ImB = -Binv
        _eye_add_(ImB, 1.0)
        H = L.gemm(Linv, ImB, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)
        scale = (inv_sigma * gbar).contiguous()
        av = (_matTvec(Linv, v1) * scale[:, None]).contiguous()
        dKuf = L.gemm(H, A, alpha_vec=scale, rowvec=av, colvec=w)
        S = B + Binv + v1[:, :, None] * v1[:, None, :]
        _eye_add_(S, -2.0)                                   # S = Abar A^T = B - 2I + B^-1 + v v^T   (B = AAT + I)
        U = L.gemm(S, Linv, flags=L.GEMM_B_LOWER)
        dKuu = L.gemm(Linv, U, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR,
                      alpha_vec=(-0.5 * gbar).contiguous())
        trS = S.diagonal(dim1=1, dim2=2).sum(1)
        dnoise = (-0.5 * N / noise + 0.5 * yy / noise ** 2 + 0.5 * sum_kdiag / noise ** 2
                  - (trS + (Atv * u).sum(1)) / (2.0 * noise))
        return dKuf, dKuu, gbar * (-0.5 / noise), None, gbar * dnoise


class VarExp(torch.autograd.Function):
    """sum_n MpdLik.variational_expectations (gpitch/likelihoods.py:33-68,422-447).  Fmu, Fvar [W,2P,N]; Y [W,N];
    noise [W] -> [W]."""

    @staticmethod
    def forward(ctx, Fmu, Fvar, Y, noise, nlin):
        ve, dFmu, dFvar, dn, _ = L.varexp(Fmu.contiguous(), Fvar.contiguous(), Y.contiguous(), noise.contiguous(), nlin)
        ctx.save_for_backward(dFmu, dFvar, dn)
        return ve

    @staticmethod
    def backward(ctx, g):
        dFmu, dFvar, dn = ctx.saved_tensors
        return dFmu * g[:, None, None], dFvar * g[:, None, None], None, dn * g, None


class GaussKLWhite(torch.autograd.Function):
    """gauss_kl(q_mu, q_sqrt) with K=None (gpitch/pdgp.py:120-121).  q_mu [b,M], q_sqrt [b,M,M] -> [b]."""

    @staticmethod
    def forward(ctx, q_mu, q_sqrt):
        kl, dmu, dLq = L.gauss_kl_white(q_mu.contiguous(), q_sqrt.contiguous())
        ctx.save_for_backward(dmu, dLq)
        return kl

    @staticmethod
    def backward(ctx, g):
        dmu, dLq = ctx.saved_tensors
        return dmu * g[:, None], dLq * g[:, None, None]

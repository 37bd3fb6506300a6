"""Dense torch fp64 statement of the ANALYTIC forward/backward formulas the CUDA path implements
(DESIGN.md section 4), written against plain matrices so they can be checked on CPU against autograd of the
oracle (tests/test_formulas_cpu.py).  Test helper only -- mirrors the kernel sequence in csrc/ step by step."""
import numpy as np
import torch

DT = torch.float64


def tri_inv(L):
    return torch.linalg.solve_triangular(L, torch.eye(L.shape[0], dtype=DT), upper=False)


def chol_adjoint(L, Linv, Lbar):
    """Kbar = 1/2 Linv^T (P + P^T) Linv,  P = Phi(L^T Lbar)  (lower triangle, halved diagonal)."""
    P = torch.tril(L.t() @ Lbar)
    P = P - 0.5 * torch.diag(torch.diagonal(P))
    return 0.5 * Linv.t() @ (P + P.t()) @ Linv


def sgpr_fwd_bwd(Kuf, Kuu_jit, sumKdiag, y, s2):
    """Returns bound and dF/dKuf, dF/dKuu (symmetric, full), dF/ds2, dF/dsumKdiag."""
    M, N = Kuf.shape
    sig = torch.sqrt(s2)
    L = torch.linalg.cholesky(Kuu_jit)
    Linv = tri_inv(L)
    A = Linv @ Kuf / sig
    AAT = A @ A.t()
    B = AAT + torch.eye(M, dtype=DT)
    LB = torch.linalg.cholesky(B)
    LBinv = tri_inv(LB)
    Aerr = A @ y                      # [M]
    c = LBinv @ Aerr / sig
    bound = (-0.5 * N * np.log(2 * np.pi) - torch.sum(torch.log(torch.diagonal(LB))) - 0.5 * N * torch.log(s2)
             - 0.5 * (y @ y) / s2 + 0.5 * (c @ c) - 0.5 * sumKdiag / s2 + 0.5 * torch.trace(AAT))
    # backward
    Binv = LBinv.t() @ LBinv
    v = LBinv.t() @ c                 # = B^-1 A u,  u = y / sig
    u = y / sig
    w = u - A.t() @ v
    Abar = A - Binv @ A + torch.outer(v, w)
    dKuf = Linv.t() @ Abar / sig
    S = B - 2 * torch.eye(M, dtype=DT) + Binv + torch.outer(v, v)
    dKuu = -0.5 * Linv.t() @ S @ Linv
    ds2 = (-0.5 * N / s2 + 0.5 * (y @ y) / s2 ** 2 + 0.5 * sumKdiag / s2 ** 2
           - (torch.trace(S) + (A.t() @ v) @ u) / (2 * s2))
    dsumKdiag = -0.5 / s2
    return bound, dKuf, dKuu, ds2, dsumKdiag


def conditional_fwd(Kmn, Kmm_jit, kdiag, q_mu, Lq):
    Lm = torch.linalg.cholesky(Kmm_jit)
    Linv = tri_inv(Lm)
    A = Linv @ Kmn
    LTA = Lq.t() @ A
    fmean = A.t() @ q_mu
    fvar = kdiag - torch.sum(A * A, 0) + torch.sum(LTA * LTA, 0)
    return fmean, fvar, (Lm, Linv, A, LTA)


def conditional_bwd(saved, q_mu, Lq, mbar, vbar):
    """Given dF/dfmean = mbar [N], dF/dfvar = vbar [N]: returns dKmn, dKmm (sym), dkdiag, dq_mu, dLq (lower)."""
    Lm, Linv, A, LTA = saved
    M = A.shape[0]
    mubar = A @ mbar
    SD = (A * vbar) @ A.t()
    dLq = torch.tril(2 * SD @ Lq)
    Abar = torch.outer(q_mu, mbar) + 2 * (Lq @ LTA - A) * vbar
    dKmn = Linv.t() @ Abar
    AbarAT = torch.outer(q_mu, mubar) + 2 * (Lq @ Lq.t() - torch.eye(M, dtype=DT)) @ SD
    Lbar = -torch.tril(Linv.t() @ AbarAT)
    dKmm = chol_adjoint(Lm, Linv, Lbar)
    return dKmn, dKmm, vbar, mubar, dLq


def gauss_kl_white(q_mu, Lq):
    """KL and its gradients (whitened): returns kl, dq_mu, dLq (lower)."""
    M = q_mu.shape[0]
    d = torch.diagonal(Lq)
    kl = 0.5 * (q_mu @ q_mu - M - torch.sum(torch.log(d * d)) + torch.sum(Lq * Lq))
    return kl, q_mu.clone(), Lq - torch.diag(1.0 / d)


_GH_X, _GH_W = np.polynomial.hermite.hermgauss(20)
_GH_W = _GH_W / np.sqrt(np.pi)


def varexp_fwd_bwd(Fmu, Fvar, y, s2, P, nlin='logistic'):
    """MpdLik variational expectations summed over n, with analytic gradients.
    Fmu, Fvar [n, 2P]; returns sum(var_exp), dFmu, dFvar, ds2."""
    x = torch.as_tensor(_GH_X)
    wq = torch.as_tensor(_GH_W)
    mg, mf = Fmu[:, :P], Fmu[:, P:]
    vg, vf = Fvar[:, :P], Fvar[:, P:]
    s = torch.sqrt(2 * vg)
    X = x[None, None, :] * s[:, :, None] + mg[:, :, None]
    if nlin == 'logistic':
        sg = 1. / (1. + torch.exp(-2. * (X - np.pi)))
        dsg = 2 * sg * (1 - sg)
    elif nlin == 'softplus':
        sg = torch.log(torch.exp(X) + 1.)
        dsg = 1. / (1. + torch.exp(-X))
    else:
        sg = torch.exp(-2. * (X - np.pi) ** 2)
        dsg = sg * (-4. * (X - np.pi))
    E1 = (sg * wq).sum(-1)
    E2 = (sg * sg * wq).sum(-1)
    dE1_m = (dsg * wq).sum(-1)
    dE1_v = (dsg * wq * x).sum(-1) / s
    dE2_m = (2 * sg * dsg * wq).sum(-1)
    dE2_v = (2 * sg * dsg * wq * x).sum(-1) / s
    a = E1 * mf
    S = a.sum(1)
    Bt = (E2 * (vf + mf ** 2)).sum(1)
    C = S * S - (a * a).sum(1)
    quad = y * y - 2 * y * S + Bt + C
    ve = -0.5 * (quad / s2 + np.log(2 * np.pi) + torch.log(s2))
    da = (y[:, None] - S[:, None] + a) / s2
    dE1 = da * mf
    dE2 = -0.5 * (vf + mf ** 2) / s2
    dmf = da * E1 - E2 * mf / s2
    dvf = -0.5 * E2 / s2
    dmg = dE1 * dE1_m + dE2 * dE2_m
    dvg = dE1 * dE1_v + dE2 * dE2_v
    ds2 = (0.5 * quad / s2 ** 2 - 0.5 / s2).sum()
    return ve.sum(), torch.cat([dmg, dmf], 1), torch.cat([dvg, dvf], 1), ds2


def kernel_grads(kind, Kbar, a_pts, b_pts, var, ls, energy=None, freq=None, mode='reference'):
    """sum_{mn} Kbar[m,n] dK[m,n]/dtheta for K(a_pts, b_pts).  Returns dvar, dlen, de[Q], df[Q]."""
    za, xb = a_pts.reshape(-1, 1), b_pts.reshape(1, -1)
    if kind == 'diff_m12':
        d = za - xb + 1e-12
        r = torch.abs(d)
        E = torch.exp(-r / ls)
        cq = torch.cos(2 * np.pi * freq[:, None, None] * r[None])
        sq = torch.sin(2 * np.pi * freq[:, None, None] * r[None])
        k = (energy[:, None, None] * cq).sum(0)
        K = var * E * k
        dvar = (Kbar * K).sum() / var
        dlen = (Kbar * K * r).sum() / ls ** 2
        de = var * (Kbar[None] * E[None] * cq).sum((1, 2))
        df = -var * energy * 2 * np.pi * (Kbar[None] * E[None] * r[None] * sq).sum((1, 2))
        return dvar, dlen, de, df
    zt, xt = za / ls, xb / ls
    if mode == 'reference':
        s = (-2 * (zt * xt) + zt * zt) + xt * xt
    else:
        s = (zt - xt) ** 2
    r = torch.sqrt(s + 1e-12)
    if kind == 'matern32':
        s3 = np.sqrt(3.)
        K = var * (1 + s3 * r) * torch.exp(-s3 * r)
        dvar = (Kbar * K).sum() / var
        dlen = (Kbar * 3 * var * torch.exp(-s3 * r) * s).sum() / ls
        return dvar, dlen, None, None
    d = za - xb
    E = torch.exp(-r)
    cq = torch.cos(2 * np.pi * freq[:, None, None] * d[None])
    sq = torch.sin(2 * np.pi * freq[:, None, None] * d[None])
    k = (energy[:, None, None] * cq).sum(0)
    K = var * E * k
    dvar = (Kbar * K).sum() / var
    dlen = (Kbar * K * s / r).sum() / ls
    de = var * (Kbar[None] * E[None] * cq).sum((1, 2))
    df = -var * energy * 2 * np.pi * (Kbar[None] * E[None] * d[None] * sq).sum((1, 2))
    return dvar, dlen, de, df

"""Oracle restatement of gpitch/sgpr_ss.py.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch
from . import gpflow_ref as G
from . import kernels_ref as KR

DTYPE = torch.float64


def build_likelihood(X, Y, Z, kern, noise_var, reg=False, jitter=G.JITTER):
    """SGPRSS.build_likelihood, sgpr_ss.py:29-71.  kern: dict or list (Add).  Returns scalar bound."""
    num_inducing = Z.shape[0]
    num_data = float(Y.shape[0])
    output_dim = float(Y.shape[1])
    err = Y
    Kdiag = KR.Kdiag(kern, X)
    Kuf = KR.K(kern, Z, X)
    Kuu = KR.K(kern, Z) + torch.eye(num_inducing, dtype=DTYPE) * jitter
    L = torch.linalg.cholesky(Kuu)
    sigma = torch.sqrt(noise_var)
    A = torch.linalg.solve_triangular(L, Kuf, upper=False) / sigma
    AAT = torch.matmul(A, A.t())
    B = AAT + torch.eye(num_inducing, dtype=DTYPE)
    LB = torch.linalg.cholesky(B)
    Aerr = torch.matmul(A, err)
    c = torch.linalg.solve_triangular(LB, Aerr, upper=False) / sigma
    bound = -0.5 * num_data * output_dim * np.log(2 * np.pi)
    bound = bound + (-output_dim * torch.sum(torch.log(torch.diagonal(LB))))
    bound = bound - 0.5 * num_data * output_dim * torch.log(noise_var)
    bound = bound + (-0.5 * torch.sum(torch.square(err)) / noise_var)
    bound = bound + 0.5 * torch.sum(torch.square(c))
    bound = bound + (-0.5 * output_dim * torch.sum(Kdiag) / noise_var)
    bound = bound + 0.5 * output_dim * torch.sum(torch.diagonal(AAT))
    if reg:
        beta = 1000.
        kl = kern if isinstance(kern, (list, tuple)) else [kern]
        r = torch.abs(kl[0]['variance'])
        for k in kl[1:]:
            r = r + torch.abs(k['variance'])
        return bound + (-beta * r)
    return bound


def predict_f(X, Y, Z, kern, noise_var, Xnew, full_cov=False, jitter=G.JITTER):
    """Inherited SGPR.predict_f (separation.py:306) -> gpflow_ref.sgpr_build_predict."""
    return G.sgpr_build_predict(X, Y, Z, Xnew, lambda a, b: KR.K(kern, a, b), lambda a: KR.Kdiag(kern, a),
                                noise_var, full_cov=full_cov, jitter=jitter)


def build_predict_source(X, Y, kern_list, noise_var, Xnew, full_cov=False):
    """SGPRSS.build_predict_source, sgpr_ss.py:73-106 (dense GP per source; var uses the SUM kernel's Kdiag)."""
    mean, var = [], []
    K = KR.K(kern_list, X) + torch.eye(X.shape[0], dtype=DTYPE) * noise_var
    L = torch.linalg.cholesky(K)
    V = torch.linalg.solve_triangular(L, Y, upper=False)
    for i in range(len(kern_list)):
        Kx = KR.K(kern_list[i], X, Xnew)
        A = torch.linalg.solve_triangular(L, Kx, upper=False)
        smean = torch.matmul(A.t(), V)
        if full_cov:
            svar = KR.K(kern_list, Xnew) - torch.matmul(A.t(), A)
            svar = svar[:, :, None].repeat(1, 1, Y.shape[1])
        else:
            svar = KR.Kdiag(kern_list, Xnew) - torch.sum(torch.square(A), 0)
            svar = svar.reshape(-1, 1).repeat(1, Y.shape[1])
        mean.append(smean)
        var.append(svar)
    return mean, var

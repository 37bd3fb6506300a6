"""Oracle restatement of gpitch/likelihoods.py (live parts).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch
from . import gpflow_ref as G

DTYPE = torch.float64


def hermgauss1d(mean_g, var_g, H, nlinfun):
    """likelihoods.py:33-45."""
    gh_x, gh_w = G.hermgauss(H)
    gh_x = torch.as_tensor(gh_x.reshape(1, -1))
    gh_w = torch.as_tensor(gh_w.reshape(-1, 1) / np.sqrt(np.pi))
    shape = mean_g.shape
    X = gh_x * torch.sqrt(2. * var_g) + mean_g
    evaluations = nlinfun(X)
    E1 = torch.matmul(evaluations, gh_w).reshape(shape)
    E2 = torch.matmul(evaluations ** 2, gh_w).reshape(shape)
    return E1, E2


def log_lik_exp(Y, mean_g, var_g, mean_f, var_f, E1, E2, noise_var, K):
    """likelihoods.py:47-68 (explicit O(K^2) cross term, add_n order)."""
    A_l = K * [None]
    B_l = K * [None]
    C_l = []
    for i in range(K):
        A_l[i] = E1[i] * mean_f[i]
        B_l[i] = E2[i] * (var_f[i] + mean_f[i] ** 2)
    for i in range(K - 1):
        for j in range(i + 1, K):
            C_l.append(E1[i] * mean_f[i] * E1[j] * mean_f[j])
    A = sum(A_l[1:], A_l[0])
    B = sum(B_l[1:], B_l[0])
    if K == 1:
        C = 0. * mean_f[0]
    else:
        C = 2. * sum(C_l[1:], C_l[0])
    var_exp = -0.5 * ((1. / noise_var) * (Y ** 2 - 2. * Y * A + B + C) + np.log(2. * np.pi) + torch.log(noise_var))
    return var_exp


def mpdlik_variational_expectations(Fmu, Fvar, Y, noise_var, nlinfun, num_sources):
    """MpdLik.variational_expectations, likelihoods.py:325,422-447.  Fmu/Fvar [n, 2P] =
    [g_1..g_P | f_1..f_P]; Y [n,1] -> [n,1]."""
    P = num_sources
    mean_g_l, mean_f_l, var_g_l, var_f_l, E1, E2 = [], [], [], [], [], []
    H = 20
    for i in range(P):
        mean_g_l.append(Fmu[:, i].reshape(-1, 1))
        mean_f_l.append(Fmu[:, i + P].reshape(-1, 1))
        var_g_l.append(Fvar[:, i].reshape(-1, 1))
        var_f_l.append(Fvar[:, i + P].reshape(-1, 1))
        e1, e2 = hermgauss1d(mean_g_l[i], var_g_l[i], H, nlinfun)
        E1.append(e1)
        E2.append(e2)
    return log_lik_exp(Y, mean_g_l, var_g_l, mean_f_l, var_f_l, E1, E2, noise_var, P)


def modlik_variational_expectations(Fmu, Fvar, Y, noise_var, transfunc):
    """ModLik.variational_expectations, likelihoods.py:152-179.  NB column order is [f, g]."""
    H = 20
    gh_x, gh_w = G.hermgauss(H)
    gh_x = torch.as_tensor(gh_x.reshape(1, -1))
    gh_w = torch.as_tensor(gh_w.reshape(-1, 1) / np.sqrt(np.pi))
    mean_f, mean_g, var_f, var_g = [e.reshape(-1, 1) for e in (Fmu[:, 0], Fmu[:, 1], Fvar[:, 0], Fvar[:, 1])]
    shape = mean_g.shape
    X = gh_x * torch.sqrt(2. * var_g) + mean_g
    evaluations = transfunc(X)
    E1 = torch.matmul(evaluations, gh_w).reshape(shape)
    E2 = torch.matmul(evaluations ** 2, gh_w).reshape(shape)
    var_exp = -0.5 * ((1. / noise_var) * (Y ** 2 - 2. * Y * mean_f * E1 + (var_f + mean_f ** 2) * E2)
                      + np.log(2. * np.pi) + torch.log(noise_var))
    return var_exp


def gaussian_density(x, mu, var):
    """gpflow.densities.gaussian [GPflow-0.5, recalled]."""
    return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(var) - 0.5 * torch.square(mu - x) / var


def mpdlik_logp(F, Y, noise_var, nlinfun, num_sources):
    """MpdLik.logp, likelihoods.py:287-322."""
    P = num_sources
    mean = None
    for i in range(P):
        m = nlinfun(F[:, i]) * F[:, i + P]
        mean = m if mean is None else mean + m
    return gaussian_density(Y[:, 0], mean, noise_var).reshape(-1, 1)

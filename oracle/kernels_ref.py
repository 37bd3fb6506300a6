"""Oracle restatement of gpitch/matern12_spectral_mixture.py (+ GPflow Add / Matern32 glue).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch fp64, op-for-op in the reference's order.
A kernel is described by a plain dict so torch autograd can differentiate through it:
    {'kind': 'mercer_m12' | 'diff_m12' | 'matern32' | 'diff_m32' (gpitch/kernels.py Matern32sm / Matern32sml),
     'variance': t[], 'lengthscales': t[], 'energy': t[Q], 'frequency': t[Q]}
A list of such dicts is a GPflow ``Add`` kernel (left-fold sum, transcription.py:245).
"""
import numpy as np
import torch
from . import gpflow_ref as G

DTYPE = torch.float64


def make(kind, variance=1., lengthscales=1., energy=None, frequency=None):
    k = {'kind': kind, 'variance': G.as_t(variance), 'lengthscales': G.as_t(lengthscales)}
    if kind != 'matern32':
        k['energy'] = G.as_t(energy).reshape(-1)
        k['frequency'] = G.as_t(frequency).reshape(-1)
    return k


def phi_features(kern, X):
    """MercerMatern12sm.phi_features, matern12_spectral_mixture.py:123-133 -> [2Q, n]."""
    m = kern['energy'].shape[0]
    n = X.shape[0]
    phi_list = 2 * m * [None]
    for i in range(m):
        phi_list[i] = torch.sqrt(kern['energy'][i]) * torch.cos(2 * np.pi * kern['frequency'][i] * X)
        phi_list[i + m] = torch.sqrt(kern['energy'][i]) * torch.sin(2 * np.pi * kern['frequency'][i] * X)
    phi = torch.stack(phi_list)
    return phi.reshape(2 * m, n)


def mercer_matern12sm_K(kern, X, X2=None):
    """MercerMatern12sm.K, matern12_spectral_mixture.py:102-117."""
    r = G.euclid_dist(X, X2, kern['lengthscales'])
    phi = phi_features(kern, X)
    if X2 is None:
        k = torch.matmul(phi.t(), phi)
    else:
        phi2 = phi_features(kern, X2)
        k = torch.matmul(phi.t(), phi2)
    return kern['variance'] * torch.exp(-r) * k


def mercer_matern12sm_Kdiag(kern, X):
    """MercerMatern12sm.Kdiag, matern12_spectral_mixture.py:119-121 (left-fold reduce of energies)."""
    e = kern['energy']
    s = e[0]
    for i in range(1, e.shape[0]):
        s = s + e[i]
    var = kern['variance'] * s
    return torch.ones(X.shape[0], dtype=DTYPE) * torch.squeeze(var)


def matern12sm_K(kern, X, X2=None):
    """Matern12sm.K (difference form), matern12_spectral_mixture.py:38-56."""
    if X2 is None:
        X2 = X
    f = X[:, None, :]
    f2 = X2[None, :, :]
    r = torch.sqrt(torch.square(f - f2 + 1e-12))
    r1 = torch.sum(r / kern['lengthscales'], 2)
    r2 = torch.sum(2. * np.pi * kern['frequency'][0] * r, 2)
    k = kern['energy'][0] * torch.cos(r2)
    for i in range(1, kern['energy'].shape[0]):
        r2 = torch.sum(2. * np.pi * kern['frequency'][i] * r, 2)
        k = k + kern['energy'][i] * torch.cos(r2)
    return kern['variance'] * torch.exp(-r1) * k


def matern12sm_Kdiag(kern, X):
    """Matern12sm.Kdiag, matern12_spectral_mixture.py:58-62."""
    n = X.shape[0]
    var = torch.ones(n, dtype=DTYPE) * torch.squeeze(kern['energy'][0])
    for i in range(1, kern['energy'].shape[0]):
        var = var + torch.ones(n, dtype=DTYPE) * torch.squeeze(kern['energy'][i])
    return kern['variance'] * var


def matern32sm_K(kern, X, X2=None):
    """Matern32sm.K, gpitch/kernels.py:230-246 (and Matern32sml.K, :291-307, when 'lengthscales' has one entry per
    partial).  'energy' holds the per-partial variances; there is no global variance."""
    if X2 is None:
        X2 = X
    f = X[:, None, :]
    f2 = X2[None, :, :]
    r = torch.sqrt(torch.square(f - f2 + 1e-12))
    ls = kern['lengthscales'].reshape(-1)
    per_partial = ls.shape[0] > 1
    r1 = np.sqrt(3.) * torch.sum(r / ls[0], 2)
    r2 = torch.sum(2. * np.pi * kern['frequency'][0] * r, 2)
    k = kern['energy'][0] * (1. + r1) * torch.exp(-r1) * torch.cos(r2)
    for i in range(1, kern['energy'].shape[0]):
        if per_partial:
            r1 = np.sqrt(3.) * torch.sum(r / ls[i], 2)
        r2 = torch.sum(2. * np.pi * kern['frequency'][i] * r, 2)
        k = k + kern['energy'][i] * (1. + r1) * torch.exp(-r1) * torch.cos(r2)
    return k


def matern32sm_Kdiag(kern, X):
    """Matern32sm.Kdiag / Matern32sml.Kdiag, gpitch/kernels.py:248-252,309-313."""
    n = X.shape[0]
    var = torch.ones(n, dtype=DTYPE) * torch.squeeze(kern['energy'][0])
    for i in range(1, kern['energy'].shape[0]):
        var = var + torch.ones(n, dtype=DTYPE) * torch.squeeze(kern['energy'][i])
    return var


def K(kern, X, X2=None):
    """Dispatch; a list is GPflow Add: reduce(add, [k.K(X, X2) ...])."""
    if isinstance(kern, (list, tuple)):
        out = K(kern[0], X, X2)
        for k in kern[1:]:
            out = out + K(k, X, X2)
        return out
    kind = kern['kind']
    if kind == 'mercer_m12':
        return mercer_matern12sm_K(kern, X, X2)
    if kind == 'diff_m12':
        return matern12sm_K(kern, X, X2)
    if kind == 'matern32':
        return G.matern32_K(X, X2, kern['variance'], kern['lengthscales'])
    if kind == 'diff_m32':
        return matern32sm_K(kern, X, X2)
    raise ValueError(kind)


def Kdiag(kern, X):
    if isinstance(kern, (list, tuple)):
        out = Kdiag(kern[0], X)
        for k in kern[1:]:
            out = out + Kdiag(k, X)
        return out
    kind = kern['kind']
    if kind == 'mercer_m12':
        return mercer_matern12sm_Kdiag(kern, X)
    if kind == 'diff_m12':
        return matern12sm_Kdiag(kern, X)
    if kind == 'matern32':
        return G.stationary_Kdiag(X, kern['variance'])
    if kind == 'diff_m32':
        return matern32sm_Kdiag(kern, X)
    raise ValueError(kind)

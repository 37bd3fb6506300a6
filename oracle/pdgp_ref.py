"""Oracle restatement of gpitch/pdgp.py.  TEST INFRASTRUCTURE ONLY."""
import torch
from . import gpflow_ref as G
from . import kernels_ref as KR
from . import likelihoods_ref as LR
from . import methods_ref as MR

DTYPE = torch.float64


def build_prior_kl(za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, whiten=True,
                   jitter=G.JITTER):
    """Pdgp.build_prior_kl, pdgp.py:113-131."""
    P = len(kern_act)
    if whiten:
        kl_act = [G.gauss_kl(q_mu_act[i], q_sqrt_act[i]) for i in range(P)]
        kl_com = [G.gauss_kl(q_mu_com[i], q_sqrt_com[i]) for i in range(P)]
    else:
        kl_act, kl_com = [], []
        for i in range(P):
            k_a = KR.K(kern_act[i], za[i]) + torch.eye(za[i].shape[0], dtype=DTYPE) * jitter
            k_c = KR.K(kern_com[i], zc[i]) + torch.eye(zc[i].shape[0], dtype=DTYPE) * jitter
            kl_act.append(G.gauss_kl(q_mu_act[i], q_sqrt_act[i], k_a))
            kl_com.append(G.gauss_kl(q_mu_com[i], q_sqrt_com[i], k_c))
    return torch.sum(torch.stack(kl_act)) + torch.sum(torch.stack(kl_com))


def _cond(x, z, kern, q_mu, q_sqrt, whiten, jitter):
    return G.conditional(x, z, lambda a, b: KR.K(kern, a, b), lambda a: KR.Kdiag(kern, a), q_mu,
                         q_sqrt=q_sqrt, whiten=whiten, jitter=jitter)


def moments(x, za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, whiten=True,
            jitter=G.JITTER):
    """The conditional loop + concats of pdgp.py:139-164 -> fmean, fvar [n, 2P]."""
    P = len(kern_act)
    mean_act, var_act, mean_com, var_com = [], [], [], []
    for i in range(P):
        m, v = _cond(x, za[i], kern_act[i], q_mu_act[i], q_sqrt_act[i], whiten, jitter)
        mean_act.append(m); var_act.append(v)
        m, v = _cond(x, zc[i], kern_com[i], q_mu_com[i], q_sqrt_com[i], whiten, jitter)
        mean_com.append(m); var_com.append(v)
    fmean = torch.cat([torch.cat(mean_act, 1), torch.cat(mean_com, 1)], 1)
    fvar = torch.cat([torch.cat(var_act, 1), torch.cat(var_com, 1)], 1)
    return fmean, fvar


def build_likelihood(x, y, za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, noise_var,
                     whiten=True, nlinfun=MR.logistic_t, num_data=None, jitter=G.JITTER):
    """Pdgp.build_likelihood, pdgp.py:133-170."""
    P = len(kern_act)
    kl = build_prior_kl(za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, whiten, jitter)
    fmean, fvar = moments(x, za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, whiten, jitter)
    var_exp = LR.mpdlik_variational_expectations(fmean, fvar, y, noise_var, nlinfun, P)
    if num_data is None:
        num_data = x.shape[0]
    scale = float(num_data) / float(x.shape[0])
    return torch.sum(var_exp) * scale - kl


def predict_act_n_com(xnew, za, zc, kern_act, kern_com, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com,
                      whiten=True, nlinfun=MR.logistic_t, jitter=G.JITTER):
    """Pdgp.predict_act_n_com, pdgp.py:190-208 (predict_act / predict_com are its halves)."""
    P = len(kern_act)
    mean_a, var_a, mean_c, var_c, mean_source = [], [], [], [], []
    for i in range(P):
        m, v = _cond(xnew, za[i], kern_act[i], q_mu_act[i], q_sqrt_act[i], whiten, jitter)
        mean_a.append(m); var_a.append(v)
        m, v = _cond(xnew, zc[i], kern_com[i], q_mu_com[i], q_sqrt_com[i], whiten, jitter)
        mean_c.append(m); var_c.append(v)
        mean_source.append(nlinfun(mean_a[i]) * mean_c[i])
    return mean_a, var_a, mean_c, var_c, mean_source

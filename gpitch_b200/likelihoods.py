"""Modulated-GP likelihoods with the class names / methods of gpitch/likelihoods.py: ``MpdLik`` (:279-447) and its
single-source twin ``ModLik`` (:136-179).  y = sum_p sigma(g_p) f_p + eps.  variational_expectations runs the fused
Gauss-Hermite CUDA kernel (csrc/ops.cu); logp is a host NumPy formula (not on the hot path)."""
import numpy as np
import torch

from .methods import logistic_tf, nlin_name, nlin_torch
from .param import Param, Parameterized, transforms


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class Likelihood(Parameterized):
    pass


class Gaussian(Likelihood):
    """gpflow.likelihoods.Gaussian (SGPR's likelihood): a single positive `variance`, initial value 1.0."""
    def __init__(self):
        self.variance = Param(1.0, transforms.positive)


class MpdLik(Likelihood):
    def __init__(self, nlinfun=logistic_tf, num_sources=1):
        self.variance = Param(1., transforms.positive)
        self.nlinfun = nlinfun
        self.num_sources = num_sources

    def _nlin_np(self, x):
        return nlin_torch(nlin_name(self.nlinfun))(torch.as_tensor(np.asarray(x, dtype=np.float64))).numpy()

    def logp(self, F, Y):
        """likelihoods.py:287-322: Gaussian density of y around sum_i nlin(g_i) f_i.  F [n, 2P] = [g | f]."""
        F, Y = np.asarray(F, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        P = self.num_sources
        mean = (self._nlin_np(F[:, :P]) * F[:, P:2 * P]).sum(1)
        var = float(np.squeeze(self.variance.value))
        return (-0.5 * np.log(2 * np.pi) - 0.5 * np.log(var) - 0.5 * np.square(mean - Y[:, 0]) / var).reshape(-1, 1)

    def variational_expectations(self, Fmu, Fvar, Y):
        """likelihoods.py:325,422-447.  Fmu, Fvar [n, 2P] (cols 0..P-1 = g, P..2P-1 = f), Y [n, 1] -> [n, 1]."""
        from . import _lib
        Fmu, Fvar, Y = (np.asarray(a, dtype=np.float64) for a in (Fmu, Fvar, Y))
        n = Fmu.shape[0]
        out = _lib.varexp(_dev(Fmu.T[None]), _dev(Fvar.T[None]), _dev(Y.reshape(1, n)),
                          _dev([float(np.squeeze(self.variance.value))]), nlin_name(self.nlinfun), need_grad=False,
                          pointwise=True)
        return out[4][0].cpu().numpy().reshape(-1, 1)


class ModLik(Likelihood):
    """likelihoods.py:136-179 -- NB the column order is [f, g] here (unlike MpdLik)."""
    def __init__(self, transfunc=logistic_tf):
        self.variance = Param(1., transforms.positive)
        self.transfunc = transfunc

    def logp(self, F, Y):
        F, Y = np.asarray(F, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        sg = nlin_torch(nlin_name(self.transfunc))(torch.as_tensor(F[:, 1])).numpy()
        var = float(np.squeeze(self.variance.value))
        return (-0.5 * np.log(2 * np.pi) - 0.5 * np.log(var) - 0.5 * np.square(F[:, 0] * sg - Y[:, 0]) / var).reshape(-1, 1)

    def variational_expectations(self, Fmu, Fvar, Y):
        inner = MpdLik(self.transfunc, 1)
        inner.variance = self.variance.value
        Fmu, Fvar = np.asarray(Fmu, dtype=np.float64), np.asarray(Fvar, dtype=np.float64)
        return inner.variational_expectations(Fmu[:, ::-1], Fvar[:, ::-1], Y)

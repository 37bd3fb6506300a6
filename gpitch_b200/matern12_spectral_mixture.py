"""Matern-1/2 x cosine spectral-mixture kernels with the constructor signatures and attributes of
gpitch/matern12_spectral_mixture.py (Matern12sm :14-67, MercerMatern12sm :70-133), evaluated by the fused
CUDA builder (csrc/builder.cu)."""
import numpy as np
import torch

from .kernels import Kern, Stationary, _dev
from .param import Param, ParamList, transforms


def _sm_row(k, Q):
    q = k.num_partials
    e = np.zeros(Q); f = np.ones(Q)
    e[:q] = [float(np.squeeze(p.value)) for p in k.energy]
    f[:q] = [float(np.squeeze(p.value)) for p in k.frequency]
    return np.concatenate([[float(np.squeeze(k.variance.value)), float(np.squeeze(k.lengthscales.value))], e, f])


class Matern12sm(Kern):
    """Difference-form kernel: variance exp(-r/l) sum_q e_q cos(2 pi f_q r), r = |x - x' + 1e-12| (:38-56).
    Energies and frequencies are fixed by the constructor (:34), as in the reference."""
    kind = 'diff_m12'

    def __init__(self, input_dim, variance=1., lengthscales=None, energy=None, frequency=None, len_fixed=False):
        Kern.__init__(self, input_dim, active_dims=None)
        self.ARD = False
        self.num_partials = len(energy)
        self.energy = ParamList([Param(energy[i], transforms.positive) for i in range(self.num_partials)])
        self.frequency = ParamList([Param(frequency[i], transforms.positive) for i in range(self.num_partials)])
        self.variance = Param(variance, transforms.positive)
        self.lengthscales = Param(lengthscales, transforms.positive)
        self.vars_n_freqs_fixed(fix_energy=True, fix_freq=True)
        if len_fixed:
            self.lengthscales.fixed = True

    def num_q(self):
        return self.num_partials

    def hyper_row(self, Q=None):
        return _sm_row(self, Q or self.num_partials)

    def Kdiag(self, X, presliced=False):
        var = np.squeeze(self.energy[0].value)
        for i in range(1, self.num_partials):
            var = var + np.squeeze(self.energy[i].value)
        return np.full(np.asarray(X).shape[0], float(np.squeeze(self.variance.value) * var))

    def vars_n_freqs_fixed(self, fix_energy=True, fix_freq=True):
        for i in range(self.num_partials):
            self.energy[i].fixed = fix_energy
            self.frequency[i].fixed = fix_freq


class MercerMatern12sm(Stationary):
    """variance exp(-r) Phi(X)^T Phi(X2) with r = GPflow euclid_dist and Mercer cos/sin features (:102-133).
    Energies / frequencies are free positive Params (never fixed by the constructor, :83-94)."""
    kind = 'mercer_m12'

    def __init__(self, input_dim, energy=np.asarray([1.]), frequency=np.asarray([2 * np.pi]), variance=1.,
                 lengthscales=1., len_fixed=False):
        Stationary.__init__(self, input_dim, variance=variance, lengthscales=lengthscales, active_dims=None, ARD=False)
        self.num_partials = len(frequency)
        self.energy = ParamList([Param(energy[i], transforms.positive) for i in range(self.num_partials)])
        self.frequency = ParamList([Param(frequency[i], transforms.positive) for i in range(self.num_partials)])
        if len_fixed:
            self.lengthscales.fixed = True

    def num_q(self):
        return self.num_partials

    def hyper_row(self, Q=None):
        return _sm_row(self, Q or self.num_partials)

    def Kdiag(self, X, presliced=False):
        e = np.squeeze(self.energy[0].value)
        for i in range(1, self.num_partials):
            e = e + np.squeeze(self.energy[i].value)
        return np.full(np.asarray(X).shape[0], float(np.squeeze(self.variance.value) * e))

    def phi_features(self, X):
        """[2Q, n] feature matrix (cos rows then sin rows), computed by the CUDA feature kernel."""
        from . import _lib
        Q = self.num_partials
        hyp = _dev(self.hyper_row(Q)[None, None])
        feat = _lib.features(_dev(np.asarray(X, dtype=np.float64).reshape(1, -1)), hyp, 1, Q)
        return feat[0, 0, :2 * Q].cpu().numpy()

    def vars_n_freqs_fixed(self, fix_energy=True, fix_freq=True):
        for i in range(self.num_partials):
            self.energy[i].fixed = fix_energy
            self.frequency[i].fixed = fix_freq

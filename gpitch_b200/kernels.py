"""Kernel base classes with the GPflow-0.5 ``Kern`` protocol gpitch relies on (K, Kdiag, +), evaluated by the
fused CUDA builder.  ``Matern32`` is the activation kernel of gpitch/init_kernels.py:12; ``Add`` is what
``np.sum(kern_list)`` produces in gpitch/transcription.py:245 / gpitch/separation.py:257."""
import numpy as np
import torch

from .param import Param, ParamList, Parameterized, transforms


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class Kern(Parameterized):
    kind = None
    distance_mode = 'reference'     # GPflow Stationary.square_dist operation order; 'stable' = direct differences

    def __init__(self, input_dim, active_dims=None):
        if input_dim != 1:
            raise ValueError('gpitch kernels act on 1-D time stamps (input_dim=1)')
        self.input_dim = input_dim

    # -- packing for the CUDA builder
    def num_q(self):
        return 0

    def hyper_row(self, Q=None):
        """[variance, lengthscale, energy[Q], frequency[Q]] (energies zero-padded up to Q)."""
        raise NotImplementedError

    def components(self):
        return [self]

    def _matrix(self, X, X2, jitter=0.0):
        from .functions import KernelMatrix
        comps = self.components()
        kinds = {c.kind for c in comps}
        if len(kinds) != 1:
            raise NotImplementedError('Add of mixed kernel kinds is not on the gpitch hot path')
        Q = max(c.num_q() for c in comps)
        hyp = _dev(np.stack([c.hyper_row(Q) for c in comps])[None])
        a = _dev(np.asarray(X, dtype=np.float64).reshape(1, -1))
        b = a if X2 is None else _dev(np.asarray(X2, dtype=np.float64).reshape(1, -1))
        with torch.no_grad():
            K = KernelMatrix.apply(hyp, a, b, comps[0].kind, comps[0].distance_mode, jitter, False)
        return K[0].cpu().numpy()

    def K(self, X, X2=None, presliced=False):
        """Covariance matrix [len(X), len(X2)] as a NumPy array (the reference's Kern.K under AutoFlow)."""
        return self._matrix(X, X2)

    compute_K = K

    def Kdiag(self, X, presliced=False):
        raise NotImplementedError

    def __add__(self, other):
        return Add([self, other])

    def __radd__(self, other):
        if isinstance(other, (int, float)) and other == 0:      # np.sum(list) starts from 0
            return self
        return Add([other, self])


class Add(Kern):
    """GPflow Add: K = reduce(add, [k.K ...]); nested Adds are flattened, order preserved."""

    def __init__(self, kern_list):
        Kern.__init__(self, 1)
        flat = []
        for k in kern_list:
            flat.extend(k.kern_list) if isinstance(k, Add) else flat.append(k)
        self.kern_list = ParamList(flat)

    def components(self):
        return list(self.kern_list)

    def Kdiag(self, X, presliced=False):
        out = self.kern_list[0].Kdiag(X)
        for k in list(self.kern_list)[1:]:
            out = out + k.Kdiag(X)
        return out


class Stationary(Kern):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Kern.__init__(self, input_dim, active_dims)
        self.variance = Param(variance, transforms.positive)
        self.lengthscales = Param(1.0 if lengthscales is None else lengthscales, transforms.positive)
        self.ARD = ARD

    def Kdiag(self, X, presliced=False):
        return np.full(np.asarray(X).shape[0], float(np.squeeze(self.variance.value)))


class Matern32(Stationary):
    """gpflow.kernels.Matern32: variance (1 + sqrt(3) r) exp(-sqrt(3) r)."""
    kind = 'matern32'

    def hyper_row(self, Q=None):
        return np.concatenate([[float(np.squeeze(self.variance.value)), float(np.squeeze(self.lengthscales.value))],
                               np.zeros(2 * (Q or 0))])


def _sq(p):
    return float(np.squeeze(p.value))


class Matern32sm(Kern):
    """gpitch/kernels.py:204-258 (legacy init_models only): sum_i variance_i (1 + r1) exp(-r1) cos(2 pi f_i r),
    r = |x - x' + 1e-12|, r1 = sqrt(3) r / lengthscales -- the difference-form builder with a Matern-3/2 envelope
    (GPX_KIND_DIFF_M32; the builder's own variance slot is 1, the per-partial variances ride in the energy slots)."""
    kind = 'diff_m32'

    def __init__(self, input_dim, num_partials, lengthscales=None, variances=None, frequencies=None):
        Kern.__init__(self, input_dim, active_dims=None)
        self.ARD = False
        self.num_partials = num_partials
        if lengthscales is None:
            lengthscales = 1.
            variances = 0.125 * np.ones((num_partials, 1))
            frequencies = 1. * (1. + np.arange(num_partials))
        self.lengthscales = Param(lengthscales, transforms.Logistic(0., 2.))
        self.variance = ParamList([Param(variances[i], transforms.Logistic(0., 0.25)) for i in range(num_partials)])
        self.frequency = ParamList([Param(frequencies[i], transforms.positive) for i in range(num_partials)])

    def num_q(self):
        return self.num_partials

    def hyper_row(self, Q=None):
        Q = Q or self.num_partials
        e = np.zeros(Q); f = np.ones(Q)
        e[:self.num_partials] = [_sq(p) for p in self.variance]
        f[:self.num_partials] = [_sq(p) for p in self.frequency]
        return np.concatenate([[1.0, _sq(self.lengthscales)], e, f])

    def Kdiag(self, X, presliced=False):
        var = _sq(self.variance[0])
        for i in range(1, self.num_partials):
            var = var + _sq(self.variance[i])
        return np.full(np.asarray(X).shape[0], var)

    def vars_n_freqs_fixed(self, fix_var=True, fix_freq=False):
        for i in range(self.num_partials):
            self.variance[i].fixed = fix_var
            self.frequency[i].fixed = fix_freq


class _Partial(object):
    """One partial of Matern32sml seen as a single-partial Matern32sm component (its own lengthscale)."""
    kind, distance_mode = 'diff_m32', 'reference'

    def __init__(self, owner, i):
        self.owner, self.i = owner, i

    def num_q(self):
        return 1

    def hyper_row(self, Q=None):
        Q = Q or 1
        e = np.zeros(Q); f = np.ones(Q)
        e[0], f[0] = _sq(self.owner.variance[self.i]), _sq(self.owner.frequency[self.i])
        return np.concatenate([[1.0, _sq(self.owner.lengthscales[self.i])], e, f])


class Matern32sml(Kern):
    """gpitch/kernels.py:261-318: Matern32sm with one lengthscale PER partial -- evaluated as the sum of num_partials
    single-partial GPX_KIND_DIFF_M32 components (the builder's component loop), no extra kernel code."""
    kind = 'diff_m32'

    def __init__(self, input_dim, num_partials, lengthscales=None, variances=None, frequencies=None):
        Kern.__init__(self, input_dim, active_dims=None)
        self.ARD = False
        self.num_partials = num_partials
        if lengthscales is None:
            lengthscales = 1. * np.ones((num_partials, 1))
            variances = 0.125 * np.ones((num_partials, 1))
            frequencies = 1. * (1. + np.arange(num_partials))
        self.lengthscales = ParamList([Param(lengthscales[i], transforms.Logistic(0., 2.)) for i in range(num_partials)])
        self.variance = ParamList([Param(variances[i], transforms.Logistic(0., 1.)) for i in range(num_partials)])
        self.frequency = ParamList([Param(frequencies[i], transforms.positive) for i in range(num_partials)])

    def num_q(self):
        return 1

    def components(self):
        return [_Partial(self, i) for i in range(self.num_partials)]

    def Kdiag(self, X, presliced=False):
        var = _sq(self.variance[0])
        for i in range(1, self.num_partials):
            var = var + _sq(self.variance[i])
        return np.full(np.asarray(X).shape[0], var)

    def vars_n_freqs_fixed(self, fix_len=False, fix_var=False, fix_freq=False):
        for i in range(self.num_partials):
            self.variance[i].fixed = fix_var
            self.frequency[i].fixed = fix_freq
            self.lengthscales[i].fixed = fix_len

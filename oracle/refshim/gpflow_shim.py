"""Minimal ``gpflow`` 0.5 API for running the gpitch reference files on the torch ``tf`` shim.
The recalled GPflow arithmetic itself lives in oracle/gpflow_ref.py ([GPflow-0.5, recalled]); this
file only provides the class protocol (Param/ParamList/transforms/Kern/Model/AutoFlow ...)."""
import functools
import types
import numpy as np
import torch
from . import tf_shim as tf
from .. import gpflow_ref as G


# ------------------------------------------------------------------ settings
class _NS(object):
    pass


settings = _NS()
settings.dtypes = _NS()
settings.dtypes.float_type = torch.float64
settings.dtypes.int_type = torch.int32
settings.numerics = _NS()
settings.numerics.jitter_level = G.JITTER


# ------------------------------------------------------------------ transforms / params
class _Identity(object):
    def forward(self, x): return x
    def backward(self, y): return np.asarray(y, dtype=np.float64)


class _Positive(object):
    def forward(self, x): return G.positive_forward(x)
    def backward(self, y): return G.positive_backward(torch.as_tensor(np.asarray(y, dtype=np.float64))).numpy()


class _Logistic(object):
    """gpflow.transforms.Logistic(a, b) [GPflow-0.5, recalled]: y = a + (b - a) / (1 + exp(-x))."""
    def __init__(self, a=0., b=1.):
        self.a, self.b = a, b

    def forward(self, x): return self.a + (self.b - self.a) * (1. / (1. + torch.exp(-x)))

    def backward(self, y):
        p = (np.asarray(y, dtype=np.float64) - self.a) / (self.b - self.a)
        return -np.log(1. / p - 1.)


transforms = _NS()
transforms.positive = _Positive()
transforms.Identity = _Identity
transforms.Logistic = _Logistic


class Param(object):
    """Holds a free-state leaf tensor; ``tensor()`` is the constrained value the graph sees."""
    def __init__(self, array, transform=None):
        self.transform = transform if transform is not None else _Identity()
        self.fixed = False
        self.set(array)

    def set(self, array):
        arr = np.atleast_1d(np.asarray(array, dtype=np.float64)).copy()
        self._shape = np.asarray(array).shape
        self.free = torch.tensor(self.transform.backward(arr), dtype=torch.float64, requires_grad=True)

    @property
    def value(self):
        return self.tensor().detach().numpy().copy()

    def tensor(self):
        t = self.transform.forward(self.free)
        if self._shape == ():
            t = t.reshape(())
        return tf.wrap(t)


class DataHolder(object):
    def __init__(self, array, on_shape_change='raise'):
        self.array = np.asarray(array, dtype=np.float64)

    def tensor(self):
        return tf.wrap(self.array)


class Parameterized(object):
    """Attribute access returns graph tensors for Param / DataHolder / MinibatchData, mimicking tf_mode."""
    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if isinstance(v, (Param, DataHolder)):
            return v.tensor()
        return v

    def __setattr__(self, name, value):
        try:
            cur = object.__getattribute__(self, name)
        except AttributeError:
            cur = None
        if isinstance(cur, Param) and not isinstance(value, (Param, DataHolder)):
            cur.set(value)
        elif isinstance(cur, DataHolder) and not isinstance(value, (Param, DataHolder)):
            cur.array = np.asarray(value, dtype=np.float64)
        else:
            object.__setattr__(self, name, value)

    def raw(self, name):
        return object.__getattribute__(self, name)

    def params(self, prefix=''):
        """Yield (qualified name, Param) over the tree, deterministic order."""
        for k in sorted(self.__dict__):
            v = self.__dict__[k]
            if isinstance(v, Param):
                yield prefix + k, v
            elif isinstance(v, ParamList):
                for i, e in enumerate(v._list):
                    if isinstance(e, Param):
                        yield '%s%s[%d]' % (prefix, k, i), e
                    elif isinstance(e, Parameterized):
                        for q in e.params('%s%s[%d].' % (prefix, k, i)):
                            yield q
            elif isinstance(v, Parameterized) and k != '_parent':
                for q in v.params(prefix + k + '.'):
                    yield q


class ParamList(Parameterized):
    def __init__(self, lst):
        object.__setattr__(self, '_list', list(lst))

    def __getitem__(self, i):
        v = self._list[i]
        if isinstance(v, Param):
            return v.tensor()
        return v

    def __len__(self): return len(self._list)

    def __iter__(self):
        for i in range(len(self._list)):
            yield self[i]

    def raw_item(self, i): return self._list[i]

    def params(self, prefix=''):
        for i, e in enumerate(self._list):
            if isinstance(e, Param):
                yield '%s[%d]' % (prefix.rstrip('.'), i), e
            elif isinstance(e, Parameterized):
                for q in e.params('%s[%d].' % (prefix.rstrip('.'), i)):
                    yield q


def AutoFlow(*specs):
    def deco(fn):
        @functools.wraps(fn)
        def run(self, *args):
            out = fn(self, *[tf.wrap(a) for a in args])

            def conv(o):
                if isinstance(o, (list, tuple)):
                    return [conv(e) for e in o]
                return o.detach().numpy().copy()
            return conv(out)
        return run
    return deco


param = types.ModuleType('gpflow.param')
param.Param, param.ParamList, param.transforms = Param, ParamList, transforms
param.AutoFlow, param.DataHolder, param.Parameterized = AutoFlow, DataHolder, Parameterized


# ------------------------------------------------------------------ kernels
class Kern(Parameterized):
    def __init__(self, input_dim, active_dims=None):
        self.input_dim = input_dim

    def _slice(self, X, X2):
        return X, X2

    def __add__(self, other):
        return Add([self, other])

    def __radd__(self, other):          # np.sum(list) starts from 0
        if isinstance(other, (int, float)) and other == 0:
            return self
        return Add([other, self])


class Add(Kern):
    def __init__(self, kern_list):
        Kern.__init__(self, 1)
        flat = []
        for k in kern_list:
            if isinstance(k, Add):
                flat.extend(k.kern_list._list)
            else:
                flat.append(k)
        self.kern_list = ParamList(flat)

    def K(self, X, X2=None, presliced=False):
        return functools.reduce(lambda a, b: a + b, [k.K(X, X2) for k in self.kern_list._list])

    def Kdiag(self, X, presliced=False):
        return functools.reduce(lambda a, b: a + b, [k.Kdiag(X) for k in self.kern_list._list])


class Stationary(Kern):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Kern.__init__(self, input_dim, active_dims)
        self.variance = Param(variance, transforms.positive)
        self.lengthscales = Param(1.0 if lengthscales is None else lengthscales, transforms.positive)
        self.ARD = ARD

    def square_dist(self, X, X2):
        return tf.wrap(G.square_dist(X, X2, self.lengthscales))

    def euclid_dist(self, X, X2):
        return tf.wrap(G.euclid_dist(X, X2, self.lengthscales))

    def Kdiag(self, X, presliced=False):
        return tf.fill(tf.stack([tf.shape(X)[0]]), tf.squeeze(self.variance))


class Matern32(Stationary):
    def K(self, X, X2=None, presliced=False):
        return tf.wrap(G.matern32_K(X, X2, self.variance, self.lengthscales))


kernels = types.ModuleType('gpflow.kernels')
kernels.Kern, kernels.Stationary, kernels.Matern32, kernels.Add = Kern, Stationary, Matern32, Add
kernels.Matern12 = kernels.Matern52 = None


# ------------------------------------------------------------------ likelihoods, densities, quadrature
class Likelihood(Parameterized):
    def __init__(self):
        pass


class Gaussian(Likelihood):
    def __init__(self):
        Likelihood.__init__(self)
        self.variance = Param(1.0, transforms.positive)


likelihoods = types.ModuleType('gpflow.likelihoods')
likelihoods.Likelihood, likelihoods.Gaussian = Likelihood, Gaussian
likelihoods.hermgauss = G.hermgauss

quadrature = types.ModuleType('gpflow.quadrature')
quadrature.hermgauss = G.hermgauss

densities = types.ModuleType('gpflow.densities')
densities.gaussian = lambda x, mu, var: -0.5 * np.log(2 * np.pi) - 0.5 * tf.log(var) - 0.5 * tf.square(mu - x) / var


class _Zero(object):
    def __call__(self, X):
        return tf.zeros((tf.shape(X)[0], 1))


mean_functions = types.ModuleType('gpflow.mean_functions')
mean_functions.Zero = _Zero


# ------------------------------------------------------------------ models
class Model(Parameterized):
    def __init__(self, name='model'):
        pass

    def compute_log_likelihood(self):
        return self.build_likelihood()

    def objective_and_grads(self):
        """-(build_likelihood) and its gradients wrt the FREE state of every non-fixed Param,
        by torch autograd through the reference's own graph-building code."""
        named = [(n, p) for n, p in self.params() if not p.fixed]
        for _, p in named:
            p.free.grad = None
        f = -self.build_likelihood()
        grads = torch.autograd.grad(f, [p.free for _, p in named], allow_unused=True)
        return float(f), {n: (np.zeros(p.free.shape) if g is None else g.numpy().copy())
                          for (n, p), g in zip(named, grads)}


model = types.ModuleType('gpflow.model')
model.Model = Model


class SGPR(Model):
    def __init__(self, X, Y, kern, Z, mean_function=None):
        Model.__init__(self)
        self.X = DataHolder(X, on_shape_change='pass')
        self.Y = DataHolder(Y, on_shape_change='pass')
        self.kern = kern
        self.likelihood = Gaussian()
        self.mean_function = mean_function or _Zero()
        self.Z = Param(Z)
        self.num_data = X.shape[0]
        self.num_latent = Y.shape[1]

    def build_predict(self, Xnew, full_cov=False):
        m, v = G.sgpr_build_predict(self.X, self.Y, self.Z, Xnew, lambda a, b: self.kern.K(a, b),
                                    lambda a: self.kern.Kdiag(a), self.likelihood.variance, full_cov=full_cov)
        return m + self.mean_function(Xnew), v

    @AutoFlow()
    def predict_f(self, Xnew):
        return self.build_predict(Xnew)


sgpr = types.ModuleType('gpflow.sgpr')
sgpr.SGPR = SGPR


class MinibatchData(DataHolder):
    """Full-batch only (minibatch_size == len(array)) -- enough for per-evaluation parity."""
    def __init__(self, array, minibatch_size, rng=None):
        DataHolder.__init__(self, array)
        assert minibatch_size == array.shape[0], 'shim supports full batch only'


minibatch = types.ModuleType('gpflow.minibatch')
minibatch.MinibatchData = MinibatchData


def _conditional(Xnew, X, kern, f, full_cov=False, q_sqrt=None, whiten=False):
    assert not full_cov
    m, v = G.conditional(Xnew, X, lambda a, b: kern.K(a, b), lambda a: kern.Kdiag(a), f, q_sqrt=q_sqrt,
                         whiten=whiten)
    return tf.wrap(m), tf.wrap(v)


conditionals = types.ModuleType('gpflow.conditionals')
conditionals.conditional = _conditional

kullback_leiblers = types.ModuleType('gpflow.kullback_leiblers')
kullback_leiblers.gauss_kl = lambda q_mu, q_sqrt, K=None: tf.wrap(G.gauss_kl(q_mu, q_sqrt, K))

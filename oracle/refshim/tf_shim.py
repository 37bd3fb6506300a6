"""Minimal ``tensorflow`` 1.2 API on torch fp64 -- just the ops the gpitch hot path calls."""
import numpy as np
import torch

float64 = torch.float64
float32 = torch.float32
int32 = torch.int32


class T(torch.Tensor):
    """torch.Tensor subclass that lets NumPy operands (np.float64 scalars, ndarrays) mix in, the way
    tf.Tensor does.  ``__array_ufunc__ = None`` makes NumPy defer to our reflected operators."""
    __array_ufunc__ = None

    @staticmethod
    def _c(o):
        if isinstance(o, np.ndarray) or isinstance(o, np.generic):
            return torch.as_tensor(np.asarray(o, dtype=np.float64))
        return o

    def __add__(self, o): return torch.Tensor.__add__(self, T._c(o))
    def __radd__(self, o): return torch.Tensor.__radd__(self, T._c(o))
    def __sub__(self, o): return torch.Tensor.__sub__(self, T._c(o))
    def __rsub__(self, o): return torch.Tensor.__rsub__(self, T._c(o))
    def __mul__(self, o): return torch.Tensor.__mul__(self, T._c(o))
    def __rmul__(self, o): return torch.Tensor.__rmul__(self, T._c(o))
    def __truediv__(self, o): return torch.Tensor.__truediv__(self, T._c(o))
    def __rtruediv__(self, o):
        # NB torch.Tensor.__rtruediv__ is reciprocal(self) * o; TF does a true (IEEE) division, and the
        # reference's distance-by-expansion amplifies that 1-ulp difference to 3e-6 in K at t = 240 s.
        o = T._c(o)
        if not isinstance(o, torch.Tensor):
            o = torch.as_tensor(o, dtype=self.dtype)
        return torch.div(o, self).as_subclass(T)


def wrap(x):
    if isinstance(x, T):
        return x
    if isinstance(x, torch.Tensor):
        return x.as_subclass(T)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x)).as_subclass(T)      # tf auto-packs a list of tensors
    return torch.as_tensor(np.asarray(x, dtype=np.float64)).as_subclass(T)


def expand_dims(x, axis): return wrap(x).unsqueeze(axis)
def sqrt(x): return torch.sqrt(wrap(x))
def square(x): return torch.square(wrap(x))
def exp(x): return torch.exp(wrap(x))
def log(x): return torch.log(wrap(x))
def cos(x): return torch.cos(wrap(x))
def sin(x): return torch.sin(wrap(x))
def abs(x): return torch.abs(wrap(x))
def add(a, b): return wrap(a) + wrap(b)
def squeeze(x): return torch.squeeze(wrap(x))
def transpose(x, perm=None):
    x = wrap(x)
    return x.t() if perm is None else x.permute(*perm)


def reduce_sum(x, axis=None):
    x = wrap(x)
    return torch.sum(x) if axis is None else torch.sum(x, axis)


def add_n(lst):
    out = wrap(lst[0])
    for e in lst[1:]:
        out = out + wrap(e)
    return out


def shape(x):
    return tuple(wrap(x).shape)


def stack(lst, axis=0):
    if all(isinstance(e, (int, np.integer)) for e in lst):
        return tuple(int(e) for e in lst)          # a shape vector
    return torch.stack([wrap(e) for e in lst], axis).as_subclass(T)


def fill(dims, value):
    return (torch.ones(tuple(int(d) for d in dims), dtype=float64) * wrap(value)).as_subclass(T)


def reshape(x, shp):
    return wrap(x).reshape(tuple(int(s) for s in shp))


def concat(lst, axis):
    return torch.cat([wrap(e) for e in lst], axis).as_subclass(T)


def tile(x, multiples):
    return wrap(x).repeat(*[int(m) for m in multiples])


def cast(x, dtype):
    if isinstance(x, (int, float, np.integer, np.floating)):
        return float(x)
    return wrap(x).to(dtype)


def eye(n, dtype=float64):
    return torch.eye(int(n), dtype=dtype).as_subclass(T)


def zeros(shp, dtype=float64):
    return torch.zeros(tuple(int(s) for s in shp), dtype=dtype).as_subclass(T)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = wrap(a), wrap(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def cholesky(x): return torch.linalg.cholesky(wrap(x))


def matrix_triangular_solve(L, B, lower=True):
    return torch.linalg.solve_triangular(wrap(L), wrap(B), upper=not lower)


def matrix_diag_part(x): return torch.diagonal(wrap(x), dim1=-2, dim2=-1)


def matrix_band_part(x, lo, hi):
    x = wrap(x)
    if lo == -1 and hi == 0:
        return torch.tril(x)
    if lo == 0 and hi == -1:
        return torch.triu(x)
    raise NotImplementedError


class _Train:
    class AdamOptimizer:
        def __init__(self, lr): self.lr = lr


train = _Train()

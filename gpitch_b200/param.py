"""GPflow-0.5-style parameter containers (host side).  Mirrors the protocol the reference relies on:
``Param(value, transform)`` with ``.fixed``, ``ParamList``, ``DataHolder``, attribute assignment that overwrites
values in place (``model.X = x``, ``kern.variance = 1.``; gpitch/separation.py:265-277), and the packed FREE state
vector the optimiser sees (``Model.get_free_state`` / ``set_state``).  Values live on the
host as float64 NumPy arrays; the models move them to the device per evaluation."""
import numpy as np


class _Identity(object):
    def forward(self, x): return x
    def backward(self, y): return y
    def dforward(self, x): return np.ones_like(x)


class _Log1pe(object):
    """transforms.positive: y = softplus(x) + lower, lower = 1e-6 ([GPflow-0.5, recalled])."""
    def __init__(self, lower=1e-6):
        self.lower = lower

    def forward(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.logaddexp(0.0, x) + self.lower

    def backward(self, y):
        y = np.asarray(y, dtype=np.float64) - self.lower
        return y + np.log(-np.expm1(-y))

    def dforward(self, x):
        return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))


class _Logistic(object):
    """transforms.Logistic(a, b): y = a + (b - a) sigmoid(x)."""
    def __init__(self, a=0., b=1.):
        self.a, self.b = a, b

    def forward(self, x):
        return self.a + (self.b - self.a) / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))

    def backward(self, y):
        p = (np.asarray(y, dtype=np.float64) - self.a) / (self.b - self.a)
        return np.log(p) - np.log1p(-p)

    def dforward(self, x):
        s = 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))
        return (self.b - self.a) * s * (1.0 - s)


class transforms(object):
    positive = _Log1pe()
    Identity = _Identity
    Logistic = _Logistic
    Log1pe = _Log1pe


class Param(object):
    def __init__(self, array, transform=None):
        self.transform = transform if transform is not None else _Identity()
        self.fixed = False
        self._array = np.array(array, dtype=np.float64)

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    @property
    def size(self):
        return self._array.size

    def assign(self, array):
        self._array = np.array(array, dtype=np.float64).reshape(self._array.shape) if np.size(array) == self._array.size \
            else np.array(array, dtype=np.float64)

    def free(self):
        return self.transform.backward(self._array).ravel()

    def set_free(self, x):
        self._array = self.transform.forward(np.asarray(x, dtype=np.float64)).reshape(self._array.shape)

    def chain(self):
        """d(constrained)/d(free), elementwise, at the current value."""
        return self.transform.dforward(self.transform.backward(self._array)).ravel()

    def __float__(self):
        return float(self._array)

    def __repr__(self):
        return 'Param(%s%s)' % (np.array2string(self._array, threshold=6), ', fixed' if self.fixed else '')


class DataHolder(object):
    def __init__(self, array, on_shape_change='raise'):
        self._array = np.array(array, dtype=np.float64)
        self.on_shape_change = on_shape_change

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    def assign(self, array):
        array = np.array(array, dtype=np.float64)
        if array.shape != self._array.shape and self.on_shape_change == 'raise':
            raise ValueError('DataHolder shape change %s -> %s' % (self._array.shape, array.shape))
        self._array = array


class Parameterized(object):
    """Assigning a number / array to an attribute that holds a Param or DataHolder overwrites its value."""

    def __setattr__(self, name, value):
        cur = self.__dict__.get(name, None)
        if isinstance(cur, (Param, DataHolder)) and not isinstance(value, (Param, DataHolder, Parameterized)):
            cur.assign(value)
        else:
            object.__setattr__(self, name, value)

    def _children(self):
        for k in sorted(self.__dict__):
            if k.startswith('_'):
                continue
            yield k, self.__dict__[k]

    def named_params(self, prefix=''):
        """Deterministic (name, Param) walk over the tree; a Param reachable twice is reported once."""
        seen = set()

        def walk(obj, pre):
            for k, v in obj._children():
                name = (pre.rstrip('.') + k) if k.startswith('[') else pre + k
                if isinstance(v, Param):
                    if id(v) not in seen:
                        seen.add(id(v))
                        yield name, v
                elif isinstance(v, Parameterized):
                    for q in walk(v, name + '.'):
                        yield q
        for q in walk(self, prefix):
            yield q

    def free_params(self):
        return [(n, p) for n, p in self.named_params() if not p.fixed]

    def get_free_state(self):
        ps = self.free_params()
        return np.concatenate([p.free() for _, p in ps]) if ps else np.zeros(0)

    def set_state(self, x):
        x = np.asarray(x, dtype=np.float64)
        off = 0
        for _, p in self.free_params():
            p.set_free(x[off:off + p.size])
            off += p.size
        assert off == x.size, 'free-state size mismatch'

    @property
    def fixed(self):
        return all(p.fixed for _, p in self.named_params())

    @fixed.setter
    def fixed(self, val):
        for _, p in self.named_params():
            p.fixed = val


class ParamList(Parameterized):
    """List of Params / Parameterized objects (gpflow.param.ParamList)."""

    def __init__(self, lst):
        object.__setattr__(self, '_list', list(lst))

    def _children(self):
        for i, v in enumerate(self._list):
            yield '[%d]' % i, v

    def __getitem__(self, i):
        return self._list[i]

    def __setitem__(self, i, value):
        cur = self._list[i]
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.assign(value)
        else:
            self._list[i] = value

    def __len__(self):
        return len(self._list)

    def __iter__(self):
        return iter(self._list)


"""Host-side optimiser loop (north star: the optimiser stays on the host).  `optimize` mirrors GPflow-0.5
Model.optimize: a string method goes to scipy.optimize.minimize with jac=True and
non-finite gradient entries zeroed; an AdamOptimizer object runs `maxiter` Adam steps on the free state with
TensorFlow's update rule."""
import numpy as np
from scipy.optimize import OptimizeResult, minimize


class AdamOptimizer(object):
    """Stand-in for tf.train.AdamOptimizer(learning_rate) as used in demos/scripts/demo-modgp.py:44-45."""
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon


def _safe(objective):
    def f(x):
        v, g = objective(x)
        g = np.asarray(g, dtype=np.float64)
        bad = ~np.isfinite(g)
        if bad.any():
            print('Warning: inf or nan in gradient: replacing with zeros')
            g = np.where(bad, 0.0, g)
        return float(v), g
    return f


def optimize(model, method='L-BFGS-B', tol=None, callback=None, maxiter=1000, **kw):
    obj = _safe(model._objective)
    x0 = model.get_free_state()
    if isinstance(method, AdamOptimizer):
        m = np.zeros_like(x0); v = np.zeros_like(x0); x = x0.copy()
        fval = None
        for t in range(1, int(maxiter) + 1):
            fval, g = obj(x)
            m = method.beta1 * m + (1 - method.beta1) * g
            v = method.beta2 * v + (1 - method.beta2) * g * g
            lr_t = method.learning_rate * np.sqrt(1 - method.beta2 ** t) / (1 - method.beta1 ** t)
            x = x - lr_t * m / (np.sqrt(v) + method.epsilon)
            if callback is not None:
                callback(x)
        model.set_state(x)
        return OptimizeResult(x=x, success=True, message='Finished iterations.', fun=fval, nit=int(maxiter))
    options = dict(maxiter=maxiter, disp=kw.pop('disp', False))
    options.update(kw.pop('options', {}))
    res = minimize(fun=obj, x0=x0, method=method, jac=True, tol=tol, callback=callback, options=options)
    model.set_state(res.x)
    return res

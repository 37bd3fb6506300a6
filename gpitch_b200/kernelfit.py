"""Kernel-profile fits of gpitch/kernelfit.py (what init_kernel(train=True) calls, gpitch/transcription.py:176-198,
gpitch/separation.py:185-207), with the profile evaluated by the CUDA builder.

``approximate_kernel`` is the Matern32sm profile k(|x|, 0) (GPX_KIND_DIFF_M32, one column), so the reference's
SciPy L-BFGS-B fit -- which differentiates the RMS loss numerically, 2 + 2m + 1 profile evaluations per gradient --
gets the loss AND its exact gradient from one builder + one builder-gradient launch (``loss_and_grad``)."""
import numpy as np
import scipy.optimize as opti
import torch

from . import _lib as L


def gabor(x, v, l, f):
    """kernelfit.py:7-8 (host)."""
    return v * np.exp(-np.abs(x) / l) * np.cos(2 * np.pi * x * f)


def func(x, *p):
    """kernelfit.py:11-16: sum of len(p) // 3 Gabor atoms (v, l, f triples), host."""
    fsum = np.zeros(np.asarray(x).size)
    for i in range(len(p) // 3):
        fsum += gabor(x, p[3 * i], p[3 * i + 1], p[3 * i + 2])
    return fsum


def learn_kernel(x, y, m):
    """kernelfit.py:19-25: scipy curve_fit of m Gabor atoms from (1, 1, i + 1) starts."""
    p0 = np.array([[1., 1., i + 1.] for i in range(m)]).reshape(-1)
    return opti.curve_fit(func, x, y, p0=p0)[0]


def _pack(p, x):
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    m = (p.size - 2) // 2
    a = np.abs(p)                                               # sqrt(p * p) in the reference
    hyp = np.concatenate([[1.0, a[1]], a[2:2 + m], a[2 + m:2 + 2 * m]])
    dev = lambda v: torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)).cuda()
    # r = |origin - pts + 1e-12| in the builder (Matern32sm.K): an origin of -1e-12 gives r = |x| as kernelfit.py uses;
    # the lags are the COLUMN points (one thread per lag)
    return p, m, dev(np.full((1, 1), -1e-12)), dev(np.abs(np.asarray(x, dtype=np.float64)).reshape(1, -1)), dev(hyp[None, None])


def approximate_kernel(p, x):
    """kernelfit.py:36-52 on the device: (1 + sqrt(3)|x|/l) exp(-sqrt(3)|x|/l) sum_i v_i cos(2 pi f_i |x|),
    p = [bias, l, v_1..v_m, f_1..f_m] (the bias enters as 0 * bias, as in the reference)."""
    p, m, origin, xs, hyp = _pack(p, x)
    K = L.kernel_build('diff_m32', 'reference', origin, xs, hyp, 1, m, None, None)
    return K[0, 0, :].cpu().numpy().reshape(np.asarray(x).shape)


def loss_and_grad(p, x, y):
    """RMS loss of kernelfit.py:28-33 and its exact gradient w.r.t. p (chain rule through sqrt(p^2) = |p|)."""
    p, m, origin, xs, hyp = _pack(p, x)
    K = L.kernel_build('diff_m32', 'reference', origin, xs, hyp, 1, m, None, None)
    yd = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1)).cuda()
    res = K[0, 0, :] - yd
    n = res.numel()
    loss = torch.sqrt((res * res).mean())
    g = np.zeros_like(p)
    if float(loss) > 0.0:
        Kbar = (res / (n * loss)).reshape(1, 1, -1).contiguous()
        dh = L.kernel_grad('diff_m32', 'reference', origin, xs, hyp, 1, m, None, None, Kbar)[0, 0].cpu().numpy()
        g[1] = dh[1]
        g[2:2 + m] = dh[2:2 + m]
        g[2 + m:2 + 2 * m] = dh[2 + m:2 + 2 * m]
        g *= np.sign(p)
    return float(loss), g


def loss_func(p, x, y):
    """kernelfit.py:28-33."""
    return loss_and_grad(p, x, y)[0]


def optimize_kern(x, y, p0, disp=False):
    """kernelfit.py:55-59: L-BFGS-B on the RMS loss (tol 1e-12) -- with the analytic gradient instead of SciPy's
    finite differences; returns |p*| like the reference."""
    phat = opti.minimize(loss_and_grad, np.asarray(p0, dtype=np.float64), jac=True, method='L-BFGS-B', args=(x, y),
                         tol=1e-12, options={'disp': disp})
    return np.sqrt(phat.x ** 2).copy()


def fit(kern, init_f, init_v, fs):
    """kernelfit.py:62-86 from the point where the initial partials are known (the reference derives init_f, init_v
    from the training audio with gpitch.init_cparam, a host-side FFT peak picker that is outside the hot path):
    returns ([lengthscale, variances, frequencies], kern_init, kern_approx)."""
    kern = np.asarray(kern, dtype=np.float64).reshape(-1)
    n = kern.size
    xkern = np.linspace(0., (n - 1.) / fs, n).reshape(-1, 1)
    p0 = np.hstack((np.array([0., 1.]), np.asarray(init_v).reshape(-1), np.asarray(init_f).reshape(-1)))
    pstar = optimize_kern(x=xkern, y=kern.reshape(-1, 1), p0=p0)
    kern_init = approximate_kernel(p0, xkern)
    kern_approx = approximate_kernel(pstar, xkern)
    npartials = (pstar.size - 2) // 2
    params = [pstar[1], pstar[2: npartials + 2], pstar[npartials + 2:]]
    return params, kern_init, kern_approx

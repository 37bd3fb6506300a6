"""Oracle restatement of gpitch/kernelfit.py (the kernel-profile fit behind init_kernel(train=True),
transcription.py:176-198 / separation.py:185-207).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy, in the
reference's operation order; Python-2 integer division restored where the reference relies on it."""
import numpy as np


def gabor(x, v, l, f):
    """kernelfit.py:7-8."""
    return v * np.exp(-np.abs(x) / l) * np.cos(2 * np.pi * x * f)


def func(x, *p):
    """kernelfit.py:11-16 (len(p)/3 is Python-2 integer division)."""
    fsum = np.zeros(x.size)
    for i in range(len(p) // 3):
        m = 3 * i
        fsum += gabor(x, p[m + 0], p[m + 1], p[m + 2])
    return fsum


def approximate_kernel(p, x):
    """kernelfit.py:36-52: Matern-3/2 envelope times a cosine mixture; p = [bias, l, v_1..v_m, f_1..f_m]."""
    nparams = p.size
    npartials = (nparams - 2) // 2
    bias = np.sqrt(p[0] * p[0])
    k_e = (1. + np.sqrt(3.) * np.abs(x) / np.sqrt(p[1] * p[1])) * np.exp(- np.sqrt(3.) *
                                                                         np.abs(x) / np.sqrt(p[1] * p[1]))
    k_partials = [np.sqrt(p[i] * p[i]) * np.cos(2 * np.pi * np.sqrt(p[i + npartials] * p[i + npartials]) * np.abs(x))
                  for i in range(2, 2 + npartials)]
    return 0. * bias + k_e * sum(k_partials)


def loss_func(p, x, y):
    """kernelfit.py:28-33: RMS error of the approximation."""
    return np.sqrt(np.square(approximate_kernel(p, x) - y).mean())

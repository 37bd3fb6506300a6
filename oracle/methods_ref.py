"""Oracle restatement of the nonlinearities in gpitch/methods.py:197-233.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch


def logistic(x):
    """methods.py:197-199 (NumPy twin)."""
    return 1. / (1. + np.exp(-2. * (x - np.pi)))


def softplus(x):
    """methods.py:205-207."""
    return np.log(np.exp(x) + 1.)


def gaussfun(x):
    """methods.py:213-214."""
    return np.exp(-2. * (x - np.pi) ** 2)


def logistic_t(x):
    """logistic_tf, methods.py:216-218."""
    return 1. / (1. + torch.exp(-2. * (x - np.pi)))


def softplus_t(x):
    """softplus_tf, methods.py:220-222."""
    return torch.log(torch.exp(x) + 1.)


def gaussfun_t(x):
    """gaussfun_tf, methods.py:232-233."""
    return torch.exp(-2. * (x - np.pi) ** 2)


NLIN = {'logistic': logistic_t, 'softplus': softplus_t, 'gauss': gaussfun_t}


def midi2freq(midi):
    """methods.py:266-267."""
    return 2. ** ((midi - 69.) / 12.) * 440.

"""Nonlinearities of gpitch/methods.py:197-233 (NumPy and torch twins) + midi2freq (:266-267)."""
import numpy as np
import torch


def logistic(x):
    """methods.py:197-199."""
    return 1. / (1. + np.exp(-2. * (x - np.pi)))


def ilogistic(x):
    """methods.py:201-203."""
    return - np.log(1. / x - 1.)


def softplus(x):
    """methods.py:205-207."""
    return np.log(np.exp(x) + 1.)


def isoftplus(x):
    """methods.py:209-211."""
    return np.log(np.exp(x) - 1.)


def gaussfun(x):
    """methods.py:213-214."""
    return np.exp(-2. * (x - np.pi) ** 2)


def logistic_tf(x):
    """methods.py:216-218 (torch tensor in/out; name kept for drop-in use as Pdgp(nlinfun=logistic_tf))."""
    return 1. / (1. + torch.exp(-2. * (x - np.pi)))


def softplus_tf(x):
    """methods.py:220-222."""
    return torch.log(torch.exp(x) + 1.)


def gaussfun_tf(x):
    """methods.py:232-233."""
    return torch.exp(-2. * (x - np.pi) ** 2)


logistic_tf.nlin_name = 'logistic'
softplus_tf.nlin_name = 'softplus'
gaussfun_tf.nlin_name = 'gauss'


def nlin_torch(name):
    return {'logistic': logistic_tf, 'softplus': softplus_tf, 'gauss': gaussfun_tf}[name]


def nlin_name(fn):
    """Map a reference-style nlinfun callable (or a name) to the kernel's enum name."""
    if isinstance(fn, str):
        return fn
    n = getattr(fn, 'nlin_name', None)
    if n is None:
        raise ValueError('nlinfun must be one of gpitch_b200.methods.{logistic_tf, softplus_tf, gaussfun_tf}')
    return n


def midi2freq(midi):
    """methods.py:266-267."""
    return 2. ** ((midi - 69.) / 12.) * 440.


def freq2midi(freq):
    """methods.py:269-270."""
    return int(69. + 12. * np.log2(freq / 440.))


def find_ideal_f0(string):
    """methods.py:26-33."""
    ideal_f0 = []
    for j in range(len(string)):
        for i in range(21, 109):
            if string[j].find('M' + str(i)) != -1:
                ideal_f0.append(midi2freq(i))
    return ideal_f0

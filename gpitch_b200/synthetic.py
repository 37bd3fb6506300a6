"""Seeded synthetic audio windows of the named benchmark shapes (generator modelled on
demos/scripts/demo-modgp.py:10-26).  NumPy only; used by bench.py, tests and __graft_entry__.smoke()."""
import numpy as np

from .methods import midi2freq

FS = 16000


def harmonic_params(midis, Q):
    """e_pq ~ 1/q^2 normalised to sum 1 (init_kernels.py:63), f_pq = q * midi2freq(p) (kept below fs/2)."""
    e = 1.0 / np.arange(1, Q + 1) ** 2
    e = e / e.sum()
    f0 = np.asarray([midi2freq(m) for m in midis])
    f = f0[:, None] * np.arange(1, Q + 1)[None, :]
    f = np.where(f < FS / 2, f, f0[:, None])
    return np.tile(e, (len(midis), 1)), f


def make_windows(W, N, midis, Q, w_offset=0, hop=None, absolute_time=True, noise_std=1e-2):
    """Returns x, y [W, N]: window w covers samples [(w_offset + w) * hop, ... + N), x = idx / fs (absolute)."""
    hop = N if hop is None else hop
    e, f = harmonic_params(midis, Q)
    x = np.empty((W, N))
    y = np.empty((W, N))
    for w in range(W):
        gw = w_offset + w
        rng = np.random.default_rng(1234 + gw)
        idx = gw * hop + np.arange(N)
        t = idx / float(FS)
        x[w] = t if absolute_time else (np.arange(N) / float(FS))
        sig = np.zeros(N)
        dur = N / float(FS)
        for p in range(len(midis)):
            c = rng.uniform(t[0], t[0] + dur, 2)
            env = np.exp(-0.5 * ((t - c[0]) / (0.15 * dur)) ** 2) + np.exp(-0.5 * ((t - c[1]) / (0.15 * dur)) ** 2)
            ph = rng.uniform(0, 2 * np.pi, Q)
            sig += env * (np.sqrt(e[p])[None, :] * np.sin(2 * np.pi * f[p][None, :] * t[:, None] + ph[None, :])).sum(1)
        sig += noise_std * rng.standard_normal(N)
        y[w] = sig / np.max(np.abs(sig))
    return x, y


def pdgp_problem(W, N, M, P, Q, w_offset=0, seed=99, act_len=1.0, act_var=3.5, com_len=0.1, midi0=60):
    """Inputs of a batched Pdgp evaluation (constrained values, window-major), as float64 NumPy arrays."""
    midis = [midi0 + i for i in range(P)]
    x, y = make_windows(W, N, midis, Q, w_offset=w_offset)
    step = N // M
    z = x[:, ::step][:, :M].copy()
    e, f = harmonic_params(midis, Q)
    out = {'x': x, 'y': y, 'za': np.tile(z[:, None, :], (1, P, 1)), 'zc': np.tile(z[:, None, :], (1, P, 1))}
    out['act_hyp'] = np.tile(np.array([act_var, act_len]), (W, P, 1))
    com = np.concatenate([np.ones((P, 1)), com_len * np.ones((P, 1)), e, f], 1)
    out['com_hyp'] = np.tile(com[None], (W, 1, 1))
    out['q_mu_act'] = np.empty((W, P, M)); out['q_mu_com'] = np.empty((W, P, M))
    out['q_sqrt_act'] = np.empty((W, P, M, M)); out['q_sqrt_com'] = np.empty((W, P, M, M))
    for w in range(W):
        rng = np.random.default_rng(seed + 7919 * (w_offset + w))
        out['q_mu_act'][w] = 0.1 * rng.standard_normal((P, M))
        out['q_mu_com'][w] = 0.1 * rng.standard_normal((P, M))
        out['q_sqrt_act'][w] = np.eye(M) + 0.01 * np.tril(rng.standard_normal((P, M, M)))
        out['q_sqrt_com'][w] = np.eye(M) + 0.01 * np.tril(rng.standard_normal((P, M, M)))
    out['noise'] = np.full(W, 1.0)
    return out


def sgpr_problem(W, N, M, P, Q, w_offset=0, com_len=0.1, midis=(60, 64, 67)):
    """Inputs of a batched SGPRSS evaluation at the reference reset point (separation.py:269-277)."""
    midis = list(midis)[:P] if P <= len(midis) else [60 + i for i in range(P)]
    x, y = make_windows(W, N, midis, Q, w_offset=w_offset)
    step = N // M
    z = x[:, ::step][:, :M].copy()
    e, f = harmonic_params(midis, Q)
    hyp = np.concatenate([np.ones((P, 1)), com_len * np.ones((P, 1)), e, f], 1)
    return {'x': x, 'y': y, 'z': z, 'hyp': np.tile(hyp[None], (W, 1, 1)), 'noise': np.full(W, 1.0)}

"""Inducing-point placement heuristics of gpitch/init_models.py:9-71 (host side, NumPy/SciPy) -- the step that
decides M per window before the hot path runs."""
import numpy as np
from scipy import signal
from scipy.signal import windows as _windows


def init_liv(x, y, num_sources=1, win_size=9, thres=0.0025, dec=1):
    """Inducing variables at the extrema (zero crossings of the smoothed gradient) of y (init_models.py:9-51).

    Faithful to a quirk of the reference: it thresholds the local energy with ``idx1 = np.where(energy > thres)``
    (a 1-tuple) and then takes ``np.argsort(idx1)``, which is ``[[0, 1, ..., k-1]]`` with k = number of extrema above
    the threshold -- so the FIRST k extrema are kept, not the k loud ones."""
    x = np.asarray(x).reshape(-1, )
    y = np.asarray(y).reshape(-1, )
    win1 = _windows.hann(1600)
    energy = signal.convolve(np.abs(y), win1, mode='same') / sum(win1)
    energy /= np.max(energy)
    win2 = _windows.hann(win_size)
    y_smooth = signal.convolve(y, win2, mode='same') / sum(win2)
    f_change_sign = np.diff(np.sign(np.gradient(y_smooth)))
    idx = np.where(f_change_sign)
    x_all, y_all, energy_all = x[idx].copy(), y[idx].copy(), energy[idx].copy()
    k = int(np.count_nonzero(energy_all > thres))
    keep = np.arange(k)                      # == np.argsort(np.where(energy_all > thres)) of the reference
    x_final = x_all[keep].copy().reshape(-1, 1)
    y_final = y_all[keep].copy().reshape(-1, 1)
    za = [x_final[::dec].copy() for _ in range(num_sources)]
    zc = [x_final[::dec].copy() for _ in range(num_sources)]
    return [za, zc], y_final[::dec]


def init_iv(x, num_sources, nivps_a, nivps_c, fs):
    """Uniform inducing variables: every fs // nivps-th sample plus the last one (init_models.py:54-71)."""
    dec_a, dec_c = int(fs) // int(nivps_a), int(fs) // int(nivps_c)
    za = [np.vstack([x[::dec_a].copy(), x[-1].copy()]) for _ in range(num_sources)]
    zc = [np.vstack([x[::dec_c].copy(), x[-1].copy()]) for _ in range(num_sources)]
    return [za, zc]


def pad_inducing(z_list, M=None):
    """Ragged inducing sets -> (Z [W, M], counts [W]).  init_liv yields a different M per window
    (gpitch/separation.py:243-246); batched kernels need one M, so short sets are padded by repeating ... nothing:
    pads are placed far outside the window (1e3 s apart from everything and from each other), where the kernel
    vanishes: Kuu gets an identity-like block (variance + jitter on the diagonal), Kuf zero rows, so the bound, its
    gradients and the predictions are unchanged up to exp(-1e3 / l) = 0."""
    counts = np.asarray([np.asarray(z).size for z in z_list])
    M = int(counts.max()) if M is None else int(M)
    Z = np.empty((len(z_list), M))
    for w, z in enumerate(z_list):
        z = np.asarray(z, dtype=np.float64).reshape(-1)
        Z[w, :z.size] = z
        far = (z.max() if z.size else 0.0) + 1e3 * (1 + np.arange(M - z.size))
        Z[w, z.size:] = far
    return Z, counts

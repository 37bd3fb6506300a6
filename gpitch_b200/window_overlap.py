"""Segmentation and overlap-add with the function names, arguments and INDEX ARITHMETIC of
gpitch/window_overlap.py (windowed :7-16, merged_mean :19-37, merged_variance :40-59, merged_x :62-74,
segmented :194-211, augmentate :213-220).  Host side, NumPy; results are bit-identical to the reference
(tests/test_host_cpu.py compares against vectors produced by the reference's own file).  The reference relies on
Python-2 integer division; every such division is an explicit floor division here.  Unlike the reference the
merge functions leave their input lists untouched."""
import numpy as np
from scipy.signal import windows as _windows


def _geometry(n, ws):
    half = (ws - 1) // 2
    return half, (n - ws) // half + 1


def windowed(x, y, ws):
    """50 %-overlap split: window i covers samples [i*half, i*half + ws), half = (ws-1)//2."""
    x, y = np.asarray(x), np.asarray(y)
    half, nw = _geometry(x.size, ws)
    idx = (np.arange(max(nw, 0)) * half)[:, None] + np.arange(ws)[None, :]
    xf, yf = x.reshape(-1), y.reshape(-1)
    return [xf[r].reshape(-1, 1) for r in idx], [yf[r].reshape(-1, 1) for r in idx]


def _hann_stack(nw, ws, half, power):
    win = np.tile(_windows.hann(ws), (nw, 1))
    win[0, :half] = 1.
    if nw > 1:
        win[-1, ws - half:] = 1.          # win[-half:] = 1
    return win ** 2 if power == 2 else win


def _overlap_add(y, ws, n, power):
    nw = len(y)
    half = (ws - 1) // 2
    Y = np.stack([np.asarray(w, dtype=np.float64).reshape(-1) for w in y]) * _hann_stack(nw, ws, half, power)
    out = np.zeros((n, 1))
    out[0:half, 0] = Y[0, :half]
    out[n - half:, 0] = Y[-1, ws - half:]
    if nw > 1:
        # iteration i of the reference writes samples (i+1)*half .. (i+2)*half inclusive; the next iteration
        # overwrites the last of them, so only the final iteration's closing sample survives.
        mid = Y[:-1, half:2 * half] + Y[1:, :half]
        out[half:nw * half, 0] = mid.reshape(-1)
        out[nw * half, 0] = Y[-2, 2 * half] + Y[-1, half]
    return out


def merged_mean(y, ws, n):
    """Hann-weighted overlap-add of per-window means."""
    return _overlap_add(y, ws, n, 1)


def merged_variance(y, ws, n):
    """Hann^2-weighted overlap-add of per-window variances."""
    return _overlap_add(y, ws, n, 2)


def merged_x(x, ws):
    half = (ws - 1) // 2
    nw = len(x)
    n = half * (nw - 1) + ws
    X = np.stack([np.asarray(w, dtype=np.float64).reshape(-1) for w in x])
    out = np.zeros((n, 1))
    out[0:half, 0] = X[0, :half]
    out[n - half - 1:, 0] = X[-1, ws - half - 1:]
    if nw > 1:
        out[half:nw * half, 0] = X[:-1, ws - half - 1:ws - 1].reshape(-1)
    return out


def augmentate(x, y, augment_size=1600):
    """Zero-pad y by augment_size samples on both sides and stretch the time vector accordingly."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    pad = np.zeros(augment_size)
    yaug = np.concatenate([pad, y.reshape(-1), pad]).reshape(-1, 1)
    alpha = augment_size / 16000.
    xaug = np.linspace(x[0] - alpha, x[-1] + alpha, x.size + 2 * augment_size).reshape(-1, 1)
    return xaug, yaug


def segmented(x, y, window_size=32000, aug=False):
    """Disjoint split into y.size // window_size windows (trailing samples dropped)."""
    x, y = np.asarray(x), np.asarray(y)
    xs, ys = [], []
    for i in range(y.size // window_size):
        sl = slice(i * window_size, (i + 1) * window_size)
        xa, ya = x[sl].copy(), y[sl].copy()
        if aug:
            xa, ya = augmentate(xa, ya)
        xs.append(xa)
        ys.append(ya)
    return xs, ys


def merged_mean_device(Y, ws, n):
    """merged_mean on the device: Y is a CUDA fp64 tensor [num_windows, ws] of per-window predictive means (what the
    batched engines return); only the merged [n] stream needs to leave the GPU.  Bit-identical to merged_mean."""
    import torch
    from . import _lib
    win = torch.as_tensor(_windows.hann(ws)).to(Y.device)
    return _lib.overlap_add(Y, win, n)


def merged_variance_device(Y, ws, n):
    """merged_variance (Hann^2 weights) on the device; see merged_mean_device."""
    import torch
    from . import _lib
    win = torch.as_tensor(_windows.hann(ws) ** 2).to(Y.device)
    return _lib.overlap_add(Y, win, n)

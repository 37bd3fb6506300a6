"""Oracle restatement of gpitch/window_overlap.py (NumPy; Python-2 integer '/' -> '//').
TEST INFRASTRUCTURE ONLY.  Unlike the reference, merged_* do not mutate their input list."""
import numpy as np
from scipy.signal import windows as _w


def _hann(ws):
    """scipy.signal.hann(ws) (symmetric) -- removed from scipy; same function lives in scipy.signal.windows."""
    return _w.hann(ws)


def windowed(x, y, ws):
    """window_overlap.py:7-16."""
    n = x.size
    l = (ws - 1) // 2
    nw = (n - ws) // l + 1
    xout, yout = [], []
    for i in range(nw):
        xout.append(x[i * l:i * l + ws].copy().reshape(-1, 1))
        yout.append(y[i * l:i * l + ws].copy().reshape(-1, 1))
    return xout, yout


def _merged(y, ws, n, power):
    nw = len(y)
    ll = (ws - 1) // 2
    yw = []
    for i in range(nw):
        win = _hann(ws).reshape(-1, 1)
        if i == 0:
            win[0:ll] = 1.
        elif i == nw - 1:
            win[-ll:] = 1.
        if power == 2:
            win = win ** 2
        yw.append(y[i] * win)
    yout = np.zeros((n, 1))
    yout[0:ll] = yw[0][0:ll]
    yout[-ll:] = yw[-1][-ll:]
    for i in range(nw - 1):
        yout[(i + 1) * ll:(i + 2) * ll + 1] = yw[i][ll:].copy() + yw[i + 1][0:ll + 1].copy()
    return yout


def merged_mean(y, ws, n):
    """window_overlap.py:19-37."""
    return _merged(y, ws, n, 1)


def merged_variance(y, ws, n):
    """window_overlap.py:40-59."""
    return _merged(y, ws, n, 2)


def merged_x(x, ws):
    """window_overlap.py:62-74."""
    l = (ws - 1) // 2
    nw = len(x)
    n = (ws - 1) // 2 * (nw - 1) + ws
    xout = np.zeros((n, 1))
    xout[0:l] = x[0][0:l]
    xout[-l - 1:] = x[-1][-l - 1:]
    for i in range(nw - 1):
        xout[(i + 1) * l:(i + 2) * l] = x[i][-l - 1:-1].copy()
    return xout


def augmentate(x, y, augment_size=1600):
    """window_overlap.py:213-220."""
    addzeros = np.zeros((augment_size, 1))
    yaug1 = np.append(addzeros, y.copy()).reshape(-1, 1)
    yaug = np.append(yaug1, addzeros).reshape(-1, 1)
    alpha = augment_size / 16000.
    xaug = np.linspace(x[0] - alpha, x[-1] + alpha, x.size + 2 * augment_size).reshape(-1, 1)
    return xaug, yaug


def segmented(x, y, window_size=32000, aug=False):
    """window_overlap.py:194-211."""
    num_windows = y.size // window_size
    xs, ys = [], []
    for i in range(num_windows):
        yaux = y[i * window_size:(i + 1) * window_size].copy()
        xaux = x[i * window_size:(i + 1) * window_size].copy()
        if aug:
            xaug, yaug = augmentate(xaux, yaux)
        else:
            xaug, yaug = xaux.copy(), yaux.copy()
        ys.append(yaug)
        xs.append(xaug)
    return xs, ys

"""Array-in front ends shaped like ``gpitch.transcription.AMT`` (gpitch/transcription.py:10-298) and
``gpitch.separation.SoSp`` (gpitch/separation.py:14-379): one SGPRSS model with the GPflow ``Add`` of one
MercerMatern12sm per pitch, fitted to every window of a track.

The reference fits the windows one after the other in a Python loop (``optimize``: transcription.py:265-298,
separation.py:279-313).  Here all windows of this rank's shard are advanced in lock-step by the batched L-BFGS driver
(driver.fit_sgpr_windows): one batched bound + gradient evaluation per line-search trial, ``matrix_var`` and the
predictions all-gathered over the ranks (distributed.py), overlap-add on the device.  Dataset plumbing of the
reference (MAPS / ss_amt file trees, h5 / pickle kernel caches, piano-roll objects, plotting) is out of scope: the
constructors take the test signal and the per-pitch kernel parameters (``self.params`` of the reference:
[lengthscales, energies, frequencies]) as arrays.
NB the reference never resets energies / frequencies between windows (reset_model only restores variance,
lengthscale and noise), so window i + 1 starts from window i's fitted partials; a lock-step batch starts every window
from the initial partials instead (SURVEY.md 8(e) caveat)."""
import numpy as np
import torch

from . import distributed, driver, init_kernels, window_overlap
from .audio import Audio
from .batched import BatchedSGPR
from .init_models import init_liv, pad_inducing


def _dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)


class _WindowedSGPR(object):
    """Shared machinery of AMT / SoSp: windows, inducing points, kernels, the sharded batched engine."""
    y_scale = 1.0          # AMT.reset_model feeds 20 * y (transcription.py:255), SoSp y itself (separation.py:267)
    z_stride = 1           # AMT keeps every 3rd extremum (transcription.py:236), SoSp all of them (separation.py:246)
    len_fixed = True       # init_kern_com(len_fixed=...) : SoSp True (separation.py:236), AMT False (transcription.py:227)

    def _setup(self, y, params, pitches, x, fs, window_size, overlap, reg, device, mode):
        self.pitches = list(pitches) if pitches is not None else list(range(len(params[0])))
        self.params = [list(params[0]), list(params[1]), list(params[2])]
        self.test_data = Audio(x=x, y=y, fs=fs, window_size=window_size, overlap=overlap)
        self.reg, self.mode = reg, mode
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.kern_pitches = init_kernels.init_kern_com(num_pitches=len(self.pitches), lengthscale=self.params[0],
                                                       energy=self.params[1], frequency=self.params[2],
                                                       len_fixed=self.len_fixed)
        nwin = len(self.test_data.Y)
        self.matrix_var = np.zeros((len(self.pitches), nwin))
        self.mean, self.var, self.smean, self.svar = [], [], [], []
        self.init_inducing()
        self.model = None          # (the reference's single SGPRSS object; here: the batched engine of this rank)

    def init_inducing(self):
        """init_liv per window (transcription.py:229-237 / separation.py:238-250); ragged M is padded far outside."""
        z, u = [], []
        for xw, yw in zip(self.test_data.X, self.test_data.Y):
            a, b = init_liv(x=xw, y=yw, num_sources=1)
            z.append(a[0][0][::self.z_stride])
            u.append(b[::self.z_stride])
        self.inducing = [z, u]

    def _initial_hyp(self, nwin):
        """reset_model: variance 1, lengthscale params[0][p], noise 1 for every window (transcription.py:252-263)."""
        P = len(self.pitches)
        Q = max(len(np.atleast_1d(e)) for e in self.params[1])
        hyp = np.zeros((P, 2 + 2 * Q))
        for p in range(P):
            e, f = np.atleast_1d(self.params[1][p]).ravel(), np.atleast_1d(self.params[2][p]).ravel()
            hyp[p, 0], hyp[p, 1] = 1.0, float(np.squeeze(self.params[0][p]))
            hyp[p, 2:2 + e.size], hyp[p, 2 + Q:2 + Q + f.size] = e, f
            hyp[p, 2 + Q + f.size:] = 1.0          # padded partials: zero energy, harmless frequency
        return np.tile(hyp[None], (nwin, 1, 1)), np.ones(nwin)

    def _engine(self, lo, hi):
        X = np.stack([np.asarray(a).reshape(-1) for a in self.test_data.X[lo:hi]])
        Y = self.y_scale * np.stack([np.asarray(a).reshape(-1) for a in self.test_data.Y[lo:hi]])
        Z, counts = pad_inducing(self.inducing[0][lo:hi], M=max(np.asarray(z).size for z in self.inducing[0]))
        self.num_inducing = counts
        return BatchedSGPR(_dev(X, self.device), _dev(Y, self.device), _dev(Z, self.device), mode=self.mode, reg=self.reg)

    def _fit(self, maxiter, nwin):
        """Lock-step L-BFGS on this rank's shard of the first `nwin` windows; returns (lo, hi, fit dict)."""
        world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
        rank = torch.distributed.get_rank() if world > 1 else 0
        lo, hi = distributed.shard_windows(nwin, world, rank)
        self.model = self._engine(lo, hi)
        hyp0, noise0 = self._initial_hyp(hi - lo)
        cols = torch.ones(hyp0.shape[2], dtype=torch.bool)
        cols[1] = not self.len_fixed
        fit = driver.fit_sgpr_windows(self.model, _dev(hyp0, self.device), _dev(noise0, self.device), maxiter=maxiter,
                                      train_cols=cols)
        self.fitted = fit
        mv = distributed.all_gather_windows(fit['matrix_var'].t().contiguous(), nwin)          # [nwin, P] on every rank
        self.matrix_var[:, :nwin] = mv.t().cpu().numpy()
        return lo, hi, fit


class AMT(_WindowedSGPR):
    """gpitch.transcription.AMT with array inputs.  ``optimize`` fills ``matrix_var`` [pitches, windows]
    (transcription.py:286-288), the activation map a piano-roll is read from."""
    y_scale, z_stride, len_fixed = 20.0, 3, False

    def __init__(self, y, params, pitches=None, x=None, fs=16000, window_size=2001, reg=False, overlap=False, device=None,
                 mode='reference'):
        self._setup(y, params, pitches, x, fs, window_size, overlap, reg, device, mode)

    def optimize(self, maxiter, disp=1, nwin=None):
        nwin = len(self.test_data.Y) if nwin is None else int(nwin)
        self._fit(maxiter, nwin)
        return self.matrix_var


class SoSp(_WindowedSGPR):
    """gpitch.separation.SoSp with array inputs: ``optimize`` fits every window and stores the per-window mixture and
    source posteriors (separation.py:305-313); ``predict_s`` overlap-adds them into ``esource`` (separation.py:341-379)."""
    y_scale, z_stride, len_fixed = 1.0, 1, True

    def __init__(self, y, params, pitches=None, x=None, fs=16000, window_size=2001, reg=False, device=None,
                 mode='reference'):
        self._setup(y, params, pitches, x, fs, window_size, True, reg, device, mode)
        self.esource = None

    def optimize(self, maxiter=1000, disp=1, nwin=None):
        nwin = len(self.test_data.Y) if nwin is None else int(nwin)
        lo, hi, fit = self._fit(maxiter, nwin)
        eng = self.model
        mf, vf = eng.predict_f_chunked(eng.x, fit['hyp'], fit['noise'])
        ms, vs = eng.predict_s_chunked(eng.x, fit['hyp'], fit['noise'])
        mf, vf = (distributed.all_gather_windows(t.contiguous(), nwin) for t in (mf, vf))
        ms, vs = (distributed.all_gather_windows(t.contiguous(), nwin) for t in (ms, vs))
        self._dev_pred = (ms, vs)
        self.mean = [m.cpu().numpy().reshape(-1, 1) for m in mf]
        self.var = [v.cpu().numpy().reshape(-1, 1) for v in vf]
        self.smean = [[m[p].cpu().numpy().reshape(-1, 1) for p in range(m.shape[0])] for m in ms]
        self.svar = [[v[p].cpu().numpy().reshape(-1, 1) for p in range(v.shape[0])] for v in vs]
        return self.matrix_var

    def predict_f(self, xnew=None):
        if xnew is not None:
            raise NotImplementedError('per-window model.predict_f(xnew): use BatchedSGPR.predict_f on self.model')
        return np.asarray(self.mean).reshape(-1, 1), np.asarray(self.var).reshape(-1, 1)

    def predict_s(self):
        """Overlap-add of the per-window source posteriors on the device (window_overlap.merged_mean / merged_variance,
        bit-exact), incl. the reference's n == 224001 special case; fills and returns ``esource``."""
        ms, vs = self._dev_pred
        ws, n = self.test_data.wsize, self.test_data.x.size
        self.esource = []
        for p in range(ms.shape[1]):
            m = window_overlap.merged_mean_device(ms[:, p, :].contiguous(), ws, n)
            v = window_overlap.merged_variance_device(vs[:, p, :].contiguous(), ws, n)
            if m.numel() == 224001:
                m, v = m[:-1], v[:-1]
            self.esource.append([m.cpu().numpy().reshape(-1, 1), v.cpu().numpy().reshape(-1, 1)])
        if n == 224001:
            self.test_data.x = self.test_data.x[0:-1].reshape(-1, 1)
            self.test_data.y = self.test_data.y[0:-1].reshape(-1, 1)
        return self.esource

"""Array-in ``Audio`` container with the attributes of gpitch/audio.py:6-37 (x, y, fs, name, wsize, X, Y).  Reading wav
files from the MAPS / ss_amt trees (``Audio.read``, gpitch/audio.py:26-28) is out of scope: pass the samples."""
import numpy as np

from . import window_overlap
from .window_overlap import segmented


class Audio(object):
    def __init__(self, path=None, filename=None, frames=-1, start=0, scaled=False, window_size=None, overlap=True,
                 x=None, y=None, fs=16000, name='unnamed'):
        if path is not None or filename is not None:
            raise NotImplementedError('dataset / wav loading is outside the hot path: pass x, y arrays')
        self.path, self.name, self.fs = None, name, int(fs)
        if y is None:                       # the reference's default object: one second of a 440 Hz cosine
            self.x = np.linspace(0., (self.fs - 1.) / self.fs, self.fs).reshape(-1, 1)
            self.y = np.cos(2 * np.pi * self.x * 440.)
        else:
            self.y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
            self.x = (np.arange(self.y.size) / float(self.fs)).reshape(-1, 1) if x is None else \
                np.asarray(x, dtype=np.float64).reshape(-1, 1)
        self.wsize = self.x.size if window_size is None else int(window_size)
        self.X, self.Y = self.windowed(overlap)

    def windowed(self, overlap):
        """gpitch/audio.py:30-37: 50 %-overlap windows (window_overlap.windowed) or plain segments."""
        if overlap:
            xwin, ywin = window_overlap.windowed(x=self.x, y=self.y, ws=self.wsize)
        else:
            xwin, ywin = segmented(x=self.x, y=self.y, window_size=self.wsize)
        self.X, self.Y = xwin, ywin
        return xwin, ywin

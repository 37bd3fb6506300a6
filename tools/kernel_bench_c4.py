"""Times the multi-component (GPflow Add of P = 88 pitch kernels) builder and gradient at the C4 shapes
(SGPRSS, N = 2001, M = 200, Q = 10): FP64-pipe-bound kernels, reported as component-evaluations/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gpitch_b200 import _lib as L

W, P, N, M, Q = 32, 88, 2001, 200, 10
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(0)
x = np.stack([np.arange(N) / 16000. + 0.125 * w for w in range(W)])
z = x[:, ::10][:, :M].copy()
hyp = np.zeros((W, P, 2 + 2 * Q))
hyp[:, :, 0] = rng.uniform(0.5, 2.0, (W, P)); hyp[:, :, 1] = rng.uniform(0.01, 0.2, (W, P))
hyp[:, :, 2:2 + Q] = rng.uniform(0.05, 1.0, (W, P, Q))
f0 = 440.0 * 2 ** ((np.arange(21, 21 + P) - 69) / 12.)
hyp[:, :, 2 + Q:] = f0[None, :, None] * np.arange(1, Q + 1)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
xd, zd, hd = dev(x), dev(z), dev(hyp)
fz, fx = L.features(zd, hd, P, Q), L.features(xd, hd, P, Q)
K = torch.empty(W, M, N, dtype=torch.float64, device='cuda')
Kbar = torch.randn(W, M, N, dtype=torch.float64, device='cuda')
cases = [
    ('builder  Kuf  Add of 88 MercerMatern12sm, Q=10', lambda: L.kernel_build('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx, out=K)),
    ('grad     Kuf  88 x (var, len, 10 e, 10 f)      ', lambda: L.kernel_grad('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx, Kbar)),
    ('grad     Kuf  88 x (var, len) only             ', lambda: L.kernel_grad('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx, Kbar, need_ef=False)),
]
from gpitch_b200.batched import grid_lags
lag = grid_lags(xd, zd)
lag = (lag[0], lag[1], 2 * N)
Kmm = torch.empty(W, M, M, dtype=torch.float64, device='cuda')
Kbmm = torch.randn(W, M, M, dtype=torch.float64, device='cuda')
cases += [
    ('grad-lag Kuf  88 x (var, len, 10 e, 10 f)      ', lambda: L.kernel_grad_lag('reference', zd, xd, hd, P, Q, Kbar, lag)),
    ('builder  Kuu  Add of 88 (M x M)                 ', lambda: L.kernel_build('mercer_m12', 'reference', zd, zd, hd, P, Q, fz, fz, jitter=1e-6, out=Kmm)),
    ('grad     Kuu  88 x (var, len, 10 e, 10 f) M x M ', lambda: L.kernel_grad('mercer_m12', 'reference', zd, zd, hd, P, Q, fz, fz, Kbmm)),
    ('features phi(x), phi(z) for 88 kernels          ', lambda: (L.features(zd, hd, P, Q), L.features(xd, hd, P, Q))),
]
evals = float(W) * M * N * P
for name, fn in cases:
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('%s %8.3f ms  %7.2f G component-evaluations/s' % (name, best, evals / best * 1e-6))

"""Times the HBM-bound kernels of one C3 window chunk in isolation (CUDA events, best of reps): the fused Kuf/Kuu
builder (Mercer Matern-1/2 SM, Q=10; Matern-3/2), its analytic gradient, the predictive-marginal column statistics
and the Gauss-Hermite quadrature.  Reports achieved GB/s on ALGORITHMIC bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gpitch_b200 import _lib as L, synthetic

W, P, N, M, Q = 17, 12, 4000, 400, 10
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pr = synthetic.pdgp_problem(W, N, M, P, Q)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
x, z = dev(pr['x']), dev(pr['zc'].reshape(W * P, M))
hc, ha = dev(pr['com_hyp'].reshape(W * P, 1, -1)), dev(pr['act_hyp'].reshape(W * P, 1, 2))
b = W * P
fz, fx = L.features(z, hc, 1, Q), L.features(x, hc, 1, Q)
K = torch.empty(b, M, N, dtype=torch.float64, device='cuda')
Kmm = torch.empty(b, M, M, dtype=torch.float64, device='cuda')
Kbar = torch.randn(b, M, N, dtype=torch.float64, device='cuda')
mu = torch.randn(b, M, dtype=torch.float64, device='cuda'); kd = torch.ones(b, dtype=torch.float64, device='cuda')
Fmu = torch.randn(W, 2 * P, N, dtype=torch.float64, device='cuda'); Fvar = torch.rand(W, 2 * P, N, dtype=torch.float64, device='cuda') + 0.1
Y = torch.randn(W, N, dtype=torch.float64, device='cuda'); nz = torch.ones(W, dtype=torch.float64, device='cuda')
MN = 8.0 * M * N * b
cs, cv = torch.randn(b, N, dtype=torch.float64, device='cuda'), torch.randn(b, N, dtype=torch.float64, device='cuda')
from gpitch_b200.batched import grid_lags
lag = grid_lags(x, z)
lag = (lag[0], lag[1], 2 * N)
cases = [
    ('builder  Kuf  MercerMatern12sm Q=10 reference', lambda: L.kernel_build('mercer_m12', 'reference', z, x, hc, 1, Q, fz, fx, out=K), MN),
    ('builder  Kuf  MercerMatern12sm Q=10 stable   ', lambda: L.kernel_build('mercer_m12', 'stable', z, x, hc, 1, Q, fz, fx, out=K), MN),
    ('builder  Kuf  Matern32 reference             ', lambda: L.kernel_build('matern32', 'reference', z, x, ha, 1, 0, None, None, out=K), MN),
    ('builder  Kuu  MercerMatern12sm (+jitter)     ', lambda: L.kernel_build('mercer_m12', 'reference', z, z, hc, 1, Q, fz, fz, jitter=1e-6, out=Kmm), 8.0 * M * M * b),
    ('features phi(X) Q=10                         ', lambda: L.features(x, hc, 1, Q), 8.0 * 20 * N * b),
    ('grad     Kuf  MercerMatern12sm (var,len,e,f) ', lambda: L.kernel_grad('mercer_m12', 'reference', z, x, hc, 1, Q, fz, fx, Kbar), MN),
    ('grad     Kuf  Matern32 (var,len)             ', lambda: L.kernel_grad('matern32', 'reference', z, x, ha, 1, 0, None, None, Kbar), MN),
    ('grad     Kuf  Mercer, fused adjoint epilogue   ', lambda: L.kernel_grad('mercer_m12', 'reference', z, x, hc, 1, Q, fz, fx, Kbar, epilogue=(2.0, cs, mu, cv)), MN),
    ('grad-lag Kuf  Mercer (z on the sample grid)    ', lambda: L.kernel_grad_lag('reference', z, x, hc, 1, Q, Kbar, lag), MN),
    ('grad-lag Kuf  Mercer, fused adjoint epilogue   ', lambda: L.kernel_grad_lag('reference', z, x, hc, 1, Q, Kbar, lag, epilogue=(2.0, cs, mu, cv)), MN),
    ('grad+z   Kuf  Mercer, fused hyper + inducing   ', lambda: L.kernel_grad('mercer_m12', 'reference', z, x, hc, 1, Q, fz, fx, Kbar, epilogue=(2.0, cs, mu, cv), with_points=True), MN),
    ('grad+z   Kuf  Matern32, fused                  ', lambda: L.kernel_grad('matern32', 'reference', z, x, ha, 1, 0, None, None, Kbar, with_points=True), MN),
    ('grad_z   Kuf  Mercer (inducing inputs)         ', lambda: L.kernel_grad_points('mercer_m12', 'reference', z, x, hc, 1, Q, fz, fx, Kbar), MN),
    ('colstats fmean,fvar from A, LTA              ', lambda: L.cond_colstats(K, Kbar, mu, kd), 2 * MN),
    ('colstats fmean,fvar from Kmn, T (mode 1)     ', lambda: L.cond_colstats(K, Kbar, mu, kd, mode=1), 2 * MN),
    ('rowdot   A mbar                              ', lambda: L.rowdot(K, Kbar[:, 0, :].contiguous()), MN),
    ('varexp   fwd+bwd P=12                        ', lambda: L.varexp(Fmu, Fvar, Y, nz, 'logistic'), 8.0 * N * W * (4 * P + 1 + 4 * P)),
]
if len(sys.argv) > 2:
    cases = [cases[int(i)] for i in sys.argv[2].split(',')]
for name, fn, nbytes in cases:
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('%s %8.3f ms  %8.1f GB/s (algorithmic)  %5.1f %% of 6553 GB/s' % (name, best, nbytes / best * 1e-6, nbytes / best * 1e-6 / 65.533))

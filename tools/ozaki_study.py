"""Accuracy study for FP64 emulation on integer tensor cores (Ozaki splitting) on the operands of the dominant product of the
C3 step, T = H A with H = L^-T (Lq Lq^T - I) and A = L^-1 Kmn, for a jitter-dominated Matern-3/2 activation group
(cond(Kmm) ~ 1e9) and a MercerMatern12sm component group.  CPU only (NumPy); the int8 x int8 -> int32 products are exact, so
emulating them in int64 is faithful.  Prints the max-norm error of T against a float128 product, for the fp64 GEMM and for
S = 5 ... 9 slices of 7 bits (row-scaled H, column-scaled A, slice pairs with s + t <= S - 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import kernels_ref as KR

torch.manual_seed(0)
DT = torch.float64
N, M = 2000, 200
x = (3.0 + torch.arange(N, dtype=DT) / 16000.).reshape(-1, 1)
z = x[::N // M][:M].clone()
rng = np.random.default_rng(1)
Lq = torch.eye(M, dtype=DT) + 0.01 * torch.tril(torch.as_tensor(rng.standard_normal((M, M))))
e = torch.as_tensor(1.0 / np.arange(1, 11) ** 2); e = e / e.sum()
f = torch.as_tensor(261.6 * np.arange(1, 11))
groups = {'activation (Matern32, l = 1.0, var = 3.5)': KR.make('matern32', 3.5, 1.0),
          'component (MercerMatern12sm, l = 0.1, Q = 10)': KR.make('mercer_m12', 1.0, 0.1, e, f)}


def split(Mx, axis, S):
    """7-bit signed slices of Mx scaled by a power of two per row (axis=1) / column (axis=0): Mx ~ 2^e sum_s q_s 2^(-7 (s + 1))."""
    mx = np.max(np.abs(Mx), axis=axis, keepdims=True)
    ex = np.ceil(np.log2(np.where(mx > 0, mx, 1.0)))
    r = Mx / 2.0 ** ex                       # |r| <= 1, exact scaling
    slices = []
    for s in range(S):
        q = np.trunc(r * 128.0)              # in [-128, 128]; |r| < 1 after the first slice keeps it within int8 in practice
        q = np.clip(q, -127, 127)
        slices.append(q.astype(np.int64))
        r = r * 128.0 - q
    return ex, slices


for name, kern in groups.items():
    Kmm = KR.K(kern, z) + 1e-6 * torch.eye(M, dtype=DT)
    Kmn = KR.K(kern, z, x)
    L = torch.linalg.cholesky(Kmm)
    Linv = torch.linalg.solve_triangular(L, torch.eye(M, dtype=DT), upper=False)
    H = (Linv.T @ (Lq @ Lq.T - torch.eye(M, dtype=DT))).numpy()
    A = (Linv @ Kmn).numpy()
    exact = (H.astype(np.longdouble) @ A.astype(np.longdouble))
    scale = np.max(np.abs(exact))
    err64 = np.max(np.abs((H @ A).astype(np.longdouble) - exact)) / scale
    print('%s: cond(Kmm) = %.1e, max |T| = %.2e' % (name, np.linalg.cond(Kmm.numpy()), float(scale)))
    print('   fp64 GEMM                 max-norm relative error %.2e' % float(err64))
    for S in (5, 6, 7, 8, 9):
        eh, hs = split(H, 1, S)
        ea, as_ = split(A, 0, S)
        acc = np.zeros(exact.shape, dtype=np.longdouble)
        nprod = 0
        for u in range(S):                   # slice pairs of equal weight share one int32 accumulator
            U = np.zeros(exact.shape, dtype=np.int64)
            for s in range(u + 1):
                U += hs[s] @ as_[u - s]
                nprod += 1
            assert np.max(np.abs(U)) < 2 ** 31
            acc += U.astype(np.longdouble) * np.longdouble(2.0) ** (-7 * (u + 2))
        T = acc * (np.longdouble(2.0) ** eh) * (np.longdouble(2.0) ** ea)
        err = np.max(np.abs(T - exact)) / scale
        print('   int8 slices S = %d (%2d MMAs) max-norm relative error %.2e' % (S, nprod, float(err)))

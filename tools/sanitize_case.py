"""One small launch of every kernel family of the library (both GEMM kernels incl. triangular / SYRK / k-weighted forms,
builder, gradients, lag-histogram gradient, Cholesky + inverse in both variants, epilogues, quadrature, KL, packing,
overlap-add) for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py` under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gpitch_b200 import _lib as L, synthetic
from gpitch_b200.batched import BatchedPdgp, BatchedSGPR, grid_lags

torch.manual_seed(0)
DT = torch.float64
b, M, N = 2, 96, 208
Lo = torch.tril(torch.randn(b, M, M, dtype=DT, device='cuda'))
X = torch.randn(b, M, N, dtype=DT, device='cuda')
w = torch.randn(b, N, dtype=DT, device='cuda')
for flags, A, B in ((0, Lo, X), (L.GEMM_A_LOWER, Lo, X), (L.GEMM_TRANS_A | L.GEMM_A_UPPER, Lo, X),
                    (L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, X, X),
                    (L.GEMM_TRANS_A | L.GEMM_TRANS_B, X, X.transpose(1, 2).contiguous().transpose(1, 2).contiguous()[:, :, :M].transpose(1, 2).contiguous())):
    L.gemm(A, B, flags=flags)
L.gemm(X, X, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=w)
L.gemm(torch.randn(b, 33, 17, dtype=DT, device='cuda'), torch.randn(b, 17, 9, dtype=DT, device='cuda'))      # cp.async kernel (odd ld)
A = X @ X.transpose(1, 2) / N + 0.1 * torch.eye(M, dtype=DT, device='cuda')
L.potrf_trinv(A.clone())
Xb = torch.randn(1, 576, 600, dtype=DT, device='cuda')
L.potrf_trinv(Xb @ Xb.transpose(1, 2) / 600 + 0.1 * torch.eye(576, dtype=DT, device='cuda'))                    # wide variant
pr = synthetic.pdgp_problem(2, 400, 40, 2, 3, act_len=0.02, com_len=0.05)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']))
e, g = eng.elbo(*[dev(pr[k]) for k in BatchedPdgp.NAMES])
assert eng._lag not in (None, False)
eng.lag_grad = False; eng._lag = None
e2, g2 = eng.elbo(*[dev(pr[k]) for k in BatchedPdgp.NAMES])
sp = synthetic.sgpr_problem(2, 400, 40, 5, 3)
s = BatchedSGPR(dev(sp['x']), dev(sp['y']), dev(sp['z']))
s.bound(dev(sp['hyp']), dev(sp['noise']))
s.predict_f(dev(sp['x']), dev(sp['hyp']), dev(sp['noise']))
s.predict_s(dev(sp['x']), dev(sp['hyp']), dev(sp['noise']))
pk = L.tril_pack(Lo); L.tril_unpack(pk, M)
L.overlap_add(torch.randn(5, 201, dtype=DT, device='cuda'), torch.rand(201, dtype=DT, device='cuda'), 100 * 4 + 201)
torch.cuda.synchronize()
print('sanitize_case ok: elbo', float(e[0]), 'kernels launched', L.launch_count(), 'tma gemm launches', L.gemm_tma_launch_count())

"""M x M x M products of one chunk (batch 204, M = 400): time per launch for the structure variants used."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpitch_b200 import _lib as L
b, M = 204, 400
Lo = torch.tril(torch.randn(b, M, M, dtype=torch.float64, device='cuda'))
D = torch.randn(b, M, M, dtype=torch.float64, device='cuda')
out = torch.empty(b, M, M, dtype=torch.float64, device='cuda')
v = torch.randn(b, M, 1, dtype=torch.float64, device='cuda')
cases = [('dense D*D', lambda: L.gemm(D, D, out=out)),
         ('Linv^T * D (A upper)', lambda: L.gemm(Lo, D, out=out, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)),
         ('D * Linv (B lower)', lambda: L.gemm(D, Lo, out=out, flags=L.GEMM_B_LOWER)),
         ('H*SD lower out', lambda: L.gemm(D, D, out=out, flags=L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER)),
         ('Lq Lq^T sym', lambda: L.gemm(Lo, Lo, out=out, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)),
         ('L^T Lbar tri-tri lower', lambda: L.gemm(Lo, Lo, out=out, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER)),
         ('GEMV Linv^T v', lambda: L.gemm(Lo, v, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER))]
tot = 0
for name, fn in cases:
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    tot += best
    print('%-28s %7.3f ms' % (name, best))
print('total %.3f ms' % tot)

"""Times the five M^2 N-class GEMM launches of one C3 window chunk (batch = 17 windows x 12 latent GPs) and a dense
reference shape; prints executed and algorithmic TFLOP/s per launch type.  Run under gpurun; ncu-friendly."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpitch_b200 import _lib as L

torch.manual_seed(0)
b, M, N = 204, 400, 4000
dev = 'cuda'
Lo = torch.tril(torch.randn(b, M, M, dtype=torch.float64, device=dev))
X = torch.randn(b, M, N, dtype=torch.float64, device=dev)
X2 = torch.randn(b, M, N, dtype=torch.float64, device=dev)
w = torch.randn(b, N, dtype=torch.float64, device=dev)
mu = torch.randn(b, M, dtype=torch.float64, device=dev)
out = torch.empty(b, M, N, dtype=torch.float64, device=dev)
outS = torch.empty(b, M, M, dtype=torch.float64, device=dev)
D = torch.randn(b, M, M, dtype=torch.float64, device=dev)

cases = [
    ('TRMM  A=Linv*Kmn        (NN, A lower)', lambda: L.gemm(Lo, X, out=out, flags=L.GEMM_A_LOWER), M * M * N * b, 0.6),
    ('TRMM  LTA=Lq^T*A        (TN, A upper)', lambda: L.gemm(Lo, X, out=out, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER), M * M * N * b, 0.6),
    ('TRMM+epilogue Abar      (NN, A lower)', lambda: L.gemm(Lo, X, out=out, flags=L.GEMM_A_LOWER, alpha=2.0, gamma=-2.0, aux=X2, colscale=w, rowvec=mu, colvec=w), M * M * N * b, 0.6),
    ('SYRK  S_D=A diag(v) A^T (NT, lower+mirror)', lambda: L.gemm(X, X, out=outS, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=w), M * M * N * b, 0.6),
    ('GEMM  dense D*X         (NN)', lambda: L.gemm(D, X, out=out), 2 * M * M * N * b, 1.0),
]
S1 = torch.randn(b, M, M, dtype=torch.float64, device=dev)
cases += [
    ('MMM   H=Linv^T*W1       (TN, A upper)', lambda: L.gemm(Lo, S1, out=outS, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER), M * M * M * b, 0.6),
    ('MMM   U=P*Linv          (NN, B lower)', lambda: L.gemm(S1, Lo, out=outS, flags=L.GEMM_B_LOWER), M * M * M * b, 0.6),
    ('MMM   W1=Lq*Lq^T        (NT, lower x upper -> lower+mirror)', lambda: L.gemm(Lo, Lo, out=outS, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER | L.GEMM_C_LOWER | L.GEMM_C_MIRROR), 2 * M * M * M * b / 3, 1.0),
    ('MMM   Lbar=H*SD         (NN, lower only)', lambda: L.gemm(S1, D, out=outS, flags=L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER), M * M * M * b, 0.6),
    ('MMM   dense             (NN)', lambda: L.gemm(S1, D, out=outS), 2 * M * M * M * b, 1.0),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
if len(sys.argv) > 2:      # comma-separated case indices (ncu captures of one launch type)
    cases = [cases[int(i)] for i in sys.argv[2].split(',')]
for name, fn, flops, exec_frac in cases:
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    alg = flops / (best * 1e-3) * 1e-12
    print('%-46s %8.3f ms  algorithmic %6.2f TFLOP/s   executed(tile-padded) %6.2f TFLOP/s' % (
        name, best, alg, alg / exec_frac if exec_frac < 1 else alg))
print('DMMA peak %.2f TFLOP/s' % L.dmma_peak(3))
print('TMA-path launches: %d of %d library launches' % (L.gemm_tma_launch_count(), L.launch_count()))

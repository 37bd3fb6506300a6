"""GPU parity of every C-ABI primitive (include/gpitch_b200.h) against the oracle / golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from oracle import gpflow_ref as G, kernels_ref as KR, likelihoods_ref as LR, methods_ref as MR
import proto_math as PM

pytestmark = pytest.mark.gpu
DT = torch.float64


def dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).cuda().contiguous()


def cpu(t):
    return t.detach().cpu()


@pytest.fixture(scope='module')
def L():
    from gpitch_b200 import _lib
    return _lib


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(80, 128, 16), (77, 133, 45), (400, 257, 400), (200, 64, 130), (128, 30, 256), (5, 3, 7)])
def test_gemm_plain(L, ta, tb, M, N, K):
    torch.manual_seed(M * 7 + N)
    batch = 3
    A = torch.randn(batch, *((K, M) if ta else (M, K)), dtype=DT, device='cuda')
    B = torch.randn(batch, *((N, K) if tb else (K, N)), dtype=DT, device='cuda')
    flags = (L.GEMM_TRANS_A if ta else 0) | (L.GEMM_TRANS_B if tb else 0)
    Cg = L.gemm(A, B, flags=flags, alpha=0.7)
    ref = 0.7 * (A.transpose(1, 2) if ta else A) @ (B.transpose(1, 2) if tb else B)
    assert relerr(cpu(Cg), cpu(ref)) < 1e-13


def _tma_used(L, fn):
    """Run fn() and report whether every gpx_gemm launch inside it took the TMA + mbarrier kernel."""
    t0, l0 = L.gemm_tma_launch_count(), L.launch_count()
    out = fn()
    return out, L.gemm_tma_launch_count() - t0


@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(400, 4000, 400), (400, 400, 400), (200, 2002, 200), (80, 64, 16), (72, 56, 40),
                                   (130, 70, 34), (2048, 256, 128), (336, 64, 64), (400, 1000, 24)])
def test_gemm_tma_path_plain(L, ta, tb, M, N, K):
    """TMA-eligible shapes (even leading dimensions): full tiles, ragged M / N / K tails, every storage form, on the
    TMA path -- asserted, not assumed."""
    torch.manual_seed(M + 3 * N + 5 * K)
    batch = 2
    A = torch.randn(batch, *((K, M) if ta else (M, K)), dtype=DT, device='cuda')
    B = torch.randn(batch, *((N, K) if tb else (K, N)), dtype=DT, device='cuda')
    flags = (L.GEMM_TRANS_A if ta else 0) | (L.GEMM_TRANS_B if tb else 0)
    Cg, used = _tma_used(L, lambda: L.gemm(A, B, flags=flags, alpha=-1.3))
    assert used == 1
    ref = -1.3 * (A.transpose(1, 2) if ta else A) @ (B.transpose(1, 2) if tb else B)
    assert relerr(cpu(Cg), cpu(ref)) < 1e-13


@pytest.mark.parametrize('M,N,K,batch', [(200, 200, 1600, 1), (400, 400, 4000, 1), (160, 160, 2000, 2), (80, 64, 1998, 3),
                                         (250, 130, 1024, 1)])
def test_gemm_split_k_over_a_cluster(L, M, N, K, batch):
    """Few output tiles and a long k-loop (the A A^T / weighted-SYRK launches of a single window): the k-tiles are split
    over the CTAs of a thread-block cluster and the partial accumulators are collected through distributed shared memory.
    Every storage form, lower-only + mirrored output, k-weights, ragged K, and bit-for-bit repeatability."""
    torch.manual_seed(M + N + K)
    for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
        A = torch.randn(batch, *((K, M) if ta else (M, K)), dtype=DT, device='cuda')
        B = torch.randn(batch, *((N, K) if tb else (K, N)), dtype=DT, device='cuda')
        flags = (L.GEMM_TRANS_A if ta else 0) | (L.GEMM_TRANS_B if tb else 0)
        Cg, used = _tma_used(L, lambda: L.gemm(A, B, flags=flags, alpha=0.5))
        ref = 0.5 * (A.transpose(1, 2) if ta else A) @ (B.transpose(1, 2) if tb else B)
        assert used == 1 and relerr(cpu(Cg), cpu(ref)) < 1e-13, (ta, tb)
        assert torch.equal(Cg, L.gemm(A, B, flags=flags, alpha=0.5))          # fixed summation order
    if M == N:
        X = torch.randn(batch, M, K, dtype=DT, device='cuda')
        w = torch.randn(batch, K, dtype=DT, device='cuda')
        C0 = torch.randn(batch, M, M, dtype=DT, device='cuda')
        S = L.gemm(X, X, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        assert relerr(cpu(S), cpu(X @ X.transpose(1, 2))) < 1e-13
        if K % 16 == 0:
            Sw = L.gemm(X, X, out=C0.clone(), flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=w, alpha=2.0, beta=1.0)
            refw = 2.0 * (X * w[:, None, :]) @ X.transpose(1, 2) + C0
            # (beta reads the lower triangle of C0; the mirrored upper triangle is written from it)
            low = torch.tril(torch.ones(M, M, dtype=torch.bool, device='cuda'))
            assert relerr(cpu(Sw[:, low]), cpu(refw[:, low])) < 1e-12


def test_gemm_tma_path_structure_and_views(L):
    """Triangular operands at MMA granularity (k-tiles straddling the diagonal), lower-only / mirrored outputs with
    pruned diagonal warp tiles, k-weights by bulk copy, shared (stride-0) operands, strided sub-matrix views, batch
    beyond the 65535 grid.z limit of the cp.async kernel."""
    torch.manual_seed(11)
    batch, M, N = 3, 400, 720
    Lo = torch.tril(torch.randn(batch, M, M, dtype=DT, device='cuda'))
    X = torch.randn(batch, M, N, dtype=DT, device='cuda')
    LoT = Lo.transpose(1, 2)
    cases = [
        ('A lower', lambda: L.gemm(Lo, X, flags=L.GEMM_A_LOWER), Lo @ X),
        ('A^T upper', lambda: L.gemm(Lo, X, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER), LoT @ X),
        ('A upper stored', lambda: L.gemm(LoT.contiguous(), X, flags=L.GEMM_A_UPPER), LoT @ X),
        ('A^T lower', lambda: L.gemm(LoT.contiguous(), X, flags=L.GEMM_TRANS_A | L.GEMM_A_LOWER), Lo @ X),
        ('B lower', lambda: L.gemm(X.transpose(1, 2).contiguous(), Lo, flags=L.GEMM_B_LOWER), X.transpose(1, 2) @ Lo),
        ('B^T upper', lambda: L.gemm(X.transpose(1, 2).contiguous(), Lo, flags=L.GEMM_TRANS_B | L.GEMM_B_UPPER),
         X.transpose(1, 2) @ LoT),
        ('lower x upper -> lower+mirror', lambda: L.gemm(Lo, Lo, flags=L.GEMM_TRANS_B | L.GEMM_A_LOWER | L.GEMM_B_UPPER |
                                                        L.GEMM_C_LOWER | L.GEMM_C_MIRROR), Lo @ LoT),
        ('upper x lower -> lower+mirror', lambda: L.gemm(Lo, Lo, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER | L.GEMM_B_LOWER |
                                                        L.GEMM_C_LOWER | L.GEMM_C_MIRROR), LoT @ Lo),
        ('lower x lower -> lower, zero upper', lambda: L.gemm(Lo, Lo, flags=L.GEMM_A_LOWER | L.GEMM_B_LOWER | L.GEMM_C_LOWER |
                                                             L.GEMM_ZERO_UPPER), Lo @ Lo),
        ('syrk', lambda: L.gemm(X, X, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR), X @ X.transpose(1, 2)),
    ]
    for name, fn, ref in cases:
        out, used = _tma_used(L, fn)
        assert used == 1, name
        assert relerr(cpu(out), cpu(ref)) < 1e-13, name
    # k-weighted SYRK (K = 720 = 45 k-tiles; weights arrive by cp.async.bulk on the stage's mbarrier)
    w = torch.randn(batch, N, dtype=DT, device='cuda')
    S, used = _tma_used(L, lambda: L.gemm(X, X, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=w))
    assert used == 1 and relerr(cpu(S), cpu((X * w[:, None, :]) @ X.transpose(1, 2))) < 1e-12
    # fused epilogue on the TMA path
    aux = torch.randn(batch, M, N, dtype=DT, device='cuda')
    cs, cv = torch.randn(batch, N, dtype=DT, device='cuda'), torch.randn(batch, N, dtype=DT, device='cuda')
    rvv, av = torch.randn(batch, M, dtype=DT, device='cuda'), torch.randn(batch, dtype=DT, device='cuda')
    C0 = torch.randn(batch, M, N, dtype=DT, device='cuda')
    out, used = _tma_used(L, lambda: L.gemm(Lo, X, out=C0.clone(), flags=L.GEMM_A_LOWER, alpha=2.0, beta=0.5, gamma=-3.0,
                                            alpha_vec=av, aux=aux, colscale=cs, rowvec=rvv, colvec=cv))
    ref = cs[:, None, :] * (2.0 * av[:, None, None] * (Lo @ X) - 3.0 * aux) + rvv[:, :, None] * cv[:, None, :] + 0.5 * C0
    assert used == 1 and relerr(cpu(out), cpu(ref)) < 1e-13
    # shared operand (batch stride 0) and strided views of a larger matrix (potrf panels use exactly these)
    out, used = _tma_used(L, lambda: L.gemm(Lo[0], X, flags=L.GEMM_A_LOWER))
    assert used == 1 and relerr(cpu(out), cpu(Lo[0][None] @ X)) < 1e-13
    big = torch.randn(batch, 464, 464, dtype=DT, device='cuda')
    Av, Bv = big[:, 64:, :64], big[:, 64:128, :64]           # [400, 64] and [64, 64] views, ld = 464
    Cv = torch.zeros(batch, 464, 464, dtype=DT, device='cuda')
    out, used = _tma_used(L, lambda: L.gemm(Av, Bv, out=Cv[:, 64:, 64:128], flags=L.GEMM_TRANS_B))
    assert used == 1 and relerr(cpu(Cv[:, 64:, 64:128]), cpu(Av @ Bv.transpose(1, 2))) < 1e-13
    assert float(Cv[:, :64].abs().max()) == 0.0 and float(Cv[:, :, 128:].abs().max()) == 0.0
    # batch > 65535 (small matrices, many latent GPs): one launch on the TMA path, sliced launches on the cp.async path
    nb = 70000
    a = torch.randn(nb, 16, 16, dtype=DT, device='cuda')
    b_ = torch.randn(nb, 16, 8, dtype=DT, device='cuda')
    out, used = _tma_used(L, lambda: L.gemm(a, b_))
    assert used == 1 and relerr(cpu(out), cpu(a @ b_)) < 1e-13
    a2 = torch.randn(nb, 9, 7, dtype=DT, device='cuda')     # odd leading dimensions: cp.async kernel, two slices
    b2 = torch.randn(nb, 7, 5, dtype=DT, device='cuda')
    out, used = _tma_used(L, lambda: L.gemm(a2, b2))
    assert used == 0 and relerr(cpu(out), cpu(a2 @ b2)) < 1e-13


def test_gemm_triangular_and_epilogue(L):
    torch.manual_seed(1)
    batch, M, N = 4, 200, 333
    Lo = torch.tril(torch.randn(batch, M, M, dtype=DT, device='cuda'))
    Bm = torch.randn(batch, M, N, dtype=DT, device='cuda')
    # A lower
    assert relerr(cpu(L.gemm(Lo, Bm, flags=L.GEMM_A_LOWER)), cpu(Lo @ Bm)) < 1e-13
    # A^T with A lower  (upper operand)
    assert relerr(cpu(L.gemm(Lo, Bm, flags=L.GEMM_TRANS_A | L.GEMM_A_UPPER)), cpu(Lo.transpose(1, 2) @ Bm)) < 1e-13
    # dense x lower
    D = torch.randn(batch, 77, M, dtype=DT, device='cuda')
    assert relerr(cpu(L.gemm(D, Lo, flags=L.GEMM_B_LOWER)), cpu(D @ Lo)) < 1e-13
    # dense x lower^T (B stored [N,K] lower => op(B) upper)
    assert relerr(cpu(L.gemm(D, Lo, flags=L.GEMM_TRANS_B | L.GEMM_B_UPPER)), cpu(D @ Lo.transpose(1, 2))) < 1e-13
    # weighted SYRK, lower + mirror
    w = torch.randn(batch, N, dtype=DT, device='cuda')
    S = L.gemm(Bm, Bm, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR, kweight=w)
    assert relerr(cpu(S), cpu((Bm * w[:, None, :]) @ Bm.transpose(1, 2))) < 1e-12
    # fused epilogue: colscale*(alpha*acc + gamma*aux) + rowvec colvec^T + beta*C, alpha_vec
    aux = torch.randn(batch, M, N, dtype=DT, device='cuda')
    cs = torch.randn(batch, N, dtype=DT, device='cuda')
    rv = torch.randn(batch, M, dtype=DT, device='cuda')
    cv = torch.randn(batch, N, dtype=DT, device='cuda')
    av = torch.rand(batch, dtype=DT, device='cuda') + 0.5
    C0 = torch.randn(batch, M, N, dtype=DT, device='cuda')
    out = C0.clone()
    L.gemm(Lo, Bm, out=out, flags=L.GEMM_A_LOWER, alpha=2.0, beta=0.5, gamma=-2.0, alpha_vec=av, aux=aux, colscale=cs,
           rowvec=rv, colvec=cv)
    ref = cs[:, None, :] * (2.0 * av[:, None, None] * (Lo @ Bm) - 2.0 * aux) + rv[:, :, None] * cv[:, None, :] + 0.5 * C0
    assert relerr(cpu(out), cpu(ref)) < 1e-13
    # lower-only output with zeroed upper
    Z = L.gemm(Lo, Lo, flags=L.GEMM_A_LOWER | L.GEMM_B_LOWER | L.GEMM_C_LOWER | L.GEMM_ZERO_UPPER)
    assert relerr(cpu(Z), cpu(torch.tril(Lo @ Lo))) < 1e-13
    # shared (unbatched) A, N = 1 GEMV, odd leading dims
    v = torch.randn(batch, M, 1, dtype=DT, device='cuda')
    assert relerr(cpu(L.gemm(Lo[0], v, flags=L.GEMM_A_LOWER, batch=batch)), cpu(Lo[0][None] @ v)) < 1e-13
    Bo = torch.randn(batch, M, N + 2, dtype=DT, device='cuda')[:, :, 1:N + 1]      # misaligned view, ld odd
    assert relerr(cpu(L.gemm(Lo, Bo, flags=L.GEMM_A_LOWER)), cpu(Lo @ Bo)) < 1e-13


# ----------------------------------------------------------------------------------------------- Cholesky
@pytest.mark.parametrize('M', [1, 20, 64, 65, 100, 200, 400, 513])
def test_potrf_trinv(L, M):
    torch.manual_seed(M)
    batch = 5
    X = torch.randn(batch, M, M + 7, dtype=DT, device='cuda')
    A = X @ X.transpose(1, 2) / M + 0.1 * torch.eye(M, dtype=DT, device='cuda')
    Lc = torch.linalg.cholesky(A)
    Lg, Linv, info = L.potrf_trinv(A.clone())
    assert int(info.abs().max()) == 0
    assert relerr(cpu(Lg), cpu(Lc)) < 1e-12
    eye = torch.eye(M, dtype=DT, device='cuda')
    assert float((Linv @ Lc - eye).abs().max()) < 1e-10
    assert float(torch.triu(Linv, 1).abs().max()) == 0.0 and float(torch.triu(Lg, 1).abs().max()) == 0.0


@pytest.mark.parametrize('M,batch', [(2048, 1), (1000, 2), (576, 3), (712, 1)])
def test_potrf_trinv_few_large_matrices(L, M, batch):
    """The right-looking factorisation + recursive block inverse taken for few, large matrices (configs[4]: M = 2048, one
    window): power-of-two and ragged block counts, a partial last diagonal block (M = 1000, 712), more than one matrix."""
    torch.manual_seed(M)
    X = torch.randn(batch, M, M + 7, dtype=DT, device='cuda')
    A = X @ X.transpose(1, 2) / M + 0.1 * torch.eye(M, dtype=DT, device='cuda')
    Lc = torch.linalg.cholesky(A)
    Lg, Linv, info = L.potrf_trinv(A.clone())
    assert int(info.abs().max()) == 0
    assert relerr(cpu(Lg), cpu(Lc)) < 1e-12
    assert float((Linv @ Lc - torch.eye(M, dtype=DT, device='cuda')).abs().max()) < 1e-9
    assert float(torch.triu(Linv, 1).abs().max()) == 0.0 and float(torch.triu(Lg, 1).abs().max()) == 0.0
    Abad = A.clone(); Abad[batch - 1, 300, 300] = -5.0
    assert L.potrf_trinv(Abad)[2].cpu().tolist() == [0] * (batch - 1) + [301]


def test_potrf_reports_non_pd(L):
    M = 100
    A = torch.eye(M, dtype=DT, device='cuda')[None].repeat(3, 1, 1)
    A[1, 70, 70] = -1.0
    _, _, info = L.potrf_trinv(A)
    assert info.cpu().tolist() == [0, 71, 0]


# ----------------------------------------------------------------------------------------------- builder
class clean_l(object):
    """Oracle with the closed-form lengthscale derivative (same forward values) -- see oracle/gpflow_ref.py."""
    def __enter__(self):
        G.CLEAN_LENGTHSCALE_GRAD = True

    def __exit__(self, *a):
        G.CLEAN_LENGTHSCALE_GRAD = False


def _hyp(var, ls, e, f):
    return dev(np.concatenate([[var, ls], np.asarray(e).ravel(), np.asarray(f).ravel()])).reshape(1, 1, -1)


@pytest.mark.parametrize('tag', ['t0', 't10', 't240'])
def test_builder_vs_reference_golden(L, tag):
    """K(Z,X), K(Z) against vectors produced by the reference's own source (oracle/make_golden.py)."""
    g = load_golden('kernels_' + tag)
    rt = lambda v: G.positive_forward(G.positive_backward(torch.as_tensor(np.asarray(v, dtype=np.float64)))).numpy()
    var, ls, e, f = rt(g['variance']), rt(g['lengthscales']), rt(g['energy']), rt(g['frequency'])
    Q = e.shape[0]
    z, x = dev(g['z'].T), dev(g['x'].T)
    hyp = _hyp(var, ls, e, f)
    fz, fx = L.features(z, hyp, 1, Q), L.features(x, hyp, 1, Q)
    Kzx = L.kernel_build('mercer_m12', 'reference', z, x, hyp, 1, Q, fz, fx)
    Kzz = L.kernel_build('mercer_m12', 'reference', z, z, hyp, 1, Q, fz, fz)
    assert relerr(cpu(Kzx[0]), g['mercer_Kzx']) < 1e-11
    assert relerr(cpu(Kzz[0]), g['mercer_Kzz']) < 1e-11
    assert relerr(cpu(fx[0, 0, :2 * Q]), g['mercer_phi']) < 1e-9      # |phase| up to 1e6 rad at t = 240 s
    Dzx = L.kernel_build('diff_m12', 'reference', z, x, hyp, 1, Q, None, None)
    assert relerr(cpu(Dzx[0]), g['diff_Kzx']) < 1e-11


@pytest.mark.parametrize('kind', ['mercer_m12', 'diff_m12', 'matern32'])
@pytest.mark.parametrize('mode', ['reference', 'stable'])
def test_builder_and_grad_vs_oracle(L, kind, mode):
    if kind == 'diff_m12' and mode == 'stable':
        pytest.skip('difference form has a single distance definition')
    rng = np.random.default_rng(7)
    W, P, Q, M, N = 3, 2, 5, 50, 301
    x = 1.0 + np.arange(W * N).reshape(W, N) / 16000.
    z = x[:, ::6][:, :M].copy()
    hyp = np.zeros((W, P, 2 + 2 * Q))
    hyp[:, :, 0] = rng.uniform(0.5, 2.0, (W, P)); hyp[:, :, 1] = rng.uniform(0.002, 0.02, (W, P))
    hyp[:, :, 2:2 + Q] = rng.uniform(0.05, 1.0, (W, P, Q)); hyp[:, :, 2 + Q:] = rng.uniform(100, 3000, (W, P, Q))
    Qk = 0 if kind == 'matern32' else Q
    hk = hyp[:, :, :2 + 2 * Qk].copy()
    zd, xd, hd = dev(z), dev(x), dev(hk)
    fz = L.features(zd, hd, P, Qk) if kind == 'mercer_m12' else None
    fx = L.features(xd, hd, P, Qk) if kind == 'mercer_m12' else None
    K = L.kernel_build(kind, mode, zd, xd, hd, P, Qk, fz, fx)
    Kbar = torch.randn(W, M, N, dtype=DT, device='cuda')
    dh = L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar)
    for w in range(W):
        ht = torch.as_tensor(hk[w]).clone().requires_grad_(True)
        kerns = [{'kind': kind, 'variance': ht[p, 0], 'lengthscales': ht[p, 1], 'energy': ht[p, 2:2 + Qk],
                  'frequency': ht[p, 2 + Qk:]} for p in range(P)]
        zt, xt = torch.as_tensor(z[w]).reshape(-1, 1), torch.as_tensor(x[w]).reshape(-1, 1)
        if mode == 'reference':
            Kref = KR.K(kerns, zt, xt)
            (Kref * cpu(Kbar[w])).sum().backward()
            gref = ht.grad
        else:   # stable distance: same kernel evaluated on origin-shifted inputs (exact differences)
            Kref = KR.K(kerns, zt - x[w, 0], xt - x[w, 0]) if kind != 'mercer_m12' else None
            if Kref is None:
                d = (zt - xt.t())
                Kref = 0
                for k in kerns:
                    r = torch.sqrt((d / k['lengthscales']) ** 2 + 1e-12)
                    Kref = Kref + k['variance'] * torch.exp(-r) * (k['energy'][:, None, None] * torch.cos(
                        2 * np.pi * k['frequency'][:, None, None] * d[None])).sum(0)
            (Kref * cpu(Kbar[w])).sum().backward()
            gref = ht.grad
        assert relerr(cpu(K[w]), Kref.detach()) < (1e-11 if mode == 'reference' else 1e-9), (w, 'K')
        got = cpu(dh[w])
        assert relerr(got[:, 0], gref[:, 0]) < 1e-10, 'dvar'
        # d/d lengthscale: fp64 autograd through the reference's distance-by-expansion is itself only good to
        # ~1e-6..1e-3 at absolute time stamps (cancellation between O(x~^2/l) terms, order-dependent inside the
        # GEMVs, irreproducible) -- see tests/test_formulas_cpu.py::test_lengthscale_grad_noise_of_reference.
        # The kernel is therefore pinned to the exact derivative (proto_math, checked against mpmath) ...
        for p_ in range(P):
            kb = cpu(Kbar[w])
            _, dl_ref, _, _ = PM.kernel_grads(kind, kb, zt[:, 0], xt[:, 0], ht[p_, 0].detach(), ht[p_, 1].detach(),
                                              ht[p_, 2:2 + Qk].detach(), ht[p_, 2 + Qk:].detach(), mode=mode)
            assert abs(float(got[p_, 1]) - float(dl_ref)) < 1e-9 * abs(float(dl_ref)), 'dlen exact'
        assert relerr(got[:, 1], gref[:, 1]) < 2e-3, 'dlen'     # ... and only loosely to the noisy autograd value
        if Qk:
            assert relerr(got[:, 2:2 + Qk], gref[:, 2:2 + Qk]) < 1e-10, 'denergy'
            assert relerr(got[:, 2 + Qk:], gref[:, 2 + Qk:]) < 1e-9, 'dfreq'


@pytest.mark.parametrize('kind', ['mercer_m12', 'matern32'])
@pytest.mark.parametrize('mode', ['reference', 'stable'])
def test_grad_points_vs_oracle(L, kind, mode):
    """d/dz of sum(Kbar * K): K(z, x) through the C ABI, K(z, z) and shared point rows (div > 1) through the autograd
    wrapper, against torch autograd through the oracle's kernels (trainable Pdgp.za / zc, gpitch/pdgp.py:80-85)."""
    from gpitch_b200.functions import KernelMatrix
    rng = np.random.default_rng(11)
    W, P, Q, M, N = 3, 2, 5, 37, 301
    x = 1.0 + np.arange(W * N).reshape(W, N) / 16000.
    z = x[:, ::8][:, :M] + rng.uniform(-2e-5, 2e-5, (W, M))
    hyp = np.zeros((W, P, 2 + 2 * Q))
    hyp[:, :, 0] = rng.uniform(0.5, 2.0, (W, P)); hyp[:, :, 1] = rng.uniform(0.002, 0.02, (W, P))
    hyp[:, :, 2:2 + Q] = rng.uniform(0.05, 1.0, (W, P, Q)); hyp[:, :, 2 + Q:] = rng.uniform(100, 3000, (W, P, Q))
    Qk = 0 if kind == 'matern32' else Q
    hk = hyp[:, :, :2 + 2 * Qk].copy()
    zd, xd, hd = dev(z), dev(x), dev(hk)
    fz = L.features(zd, hd, P, Qk) if kind == 'mercer_m12' else None
    fx = L.features(xd, hd, P, Qk) if kind == 'mercer_m12' else None
    Kbar = torch.randn(W, M, N, dtype=DT, device='cuda')
    Kbar2 = torch.randn(W, M, M, dtype=DT, device='cuda')
    dz = L.kernel_grad_points(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar)
    zl = zd.clone().requires_grad_(True)
    Kzz = KernelMatrix.apply(hd, zl, zl, kind, mode, 1e-6, True)
    (Kzz * Kbar2).sum().backward()

    def kref(kerns, a, b, shift):
        if mode == 'reference':
            return KR.K(kerns, a, b)
        if kind != 'mercer_m12':
            return KR.K(kerns, a - shift, b - shift)
        d = a - b.t()
        out = 0
        for k in kerns:
            r = torch.sqrt((d / k['lengthscales']) ** 2 + 1e-12)
            out = out + k['variance'] * torch.exp(-r) * (k['energy'][:, None, None] * torch.cos(
                2 * np.pi * k['frequency'][:, None, None] * d[None])).sum(0)
        return out

    for w in range(W):
        ht = torch.as_tensor(hk[w])
        kerns = [{'kind': kind, 'variance': ht[p, 0], 'lengthscales': ht[p, 1], 'energy': ht[p, 2:2 + Qk],
                  'frequency': ht[p, 2 + Qk:]} for p in range(P)]
        zt = torch.as_tensor(z[w]).reshape(-1, 1).clone().requires_grad_(True)
        xt = torch.as_tensor(x[w]).reshape(-1, 1)
        (kref(kerns, zt, xt, x[w, 0]) * cpu(Kbar[w])).sum().backward()
        # reference mode: the oracle's autograd through the distance-by-expansion carries ~eps * x~ / d~ noise itself
        tol = 1e-8 if mode == 'reference' else 1e-9
        assert relerr(cpu(dz[w]), zt.grad[:, 0]) < tol, (w, 'K(z,x)')
        zt2 = torch.as_tensor(z[w]).reshape(-1, 1).clone().requires_grad_(True)
        (kref(kerns, zt2, zt2, x[w, 0]) * cpu(Kbar2[w])).sum().backward()
        assert relerr(cpu(zl.grad[w]), zt2.grad[:, 0]) < tol, (w, 'K(z,z)')
    # fused adjoint epilogue (conditional()'s Kbar_mn = 2 T diag(vbar) + a mbar^T is never materialised): same result as
    # the materialised adjoint, for the hyper-parameter and the point gradients
    cs, cv = torch.randn(W, N, dtype=DT, device='cuda'), torch.randn(W, N, dtype=DT, device='cuda')
    rv = torch.randn(W, M, dtype=DT, device='cuda')
    Kmat = 2.0 * Kbar * cs[:, None, :] + rv[:, :, None] * cv[:, None, :]
    for fn in (L.kernel_grad, L.kernel_grad_points):
        fused = fn(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar, epilogue=(2.0, cs, rv, cv))
        plain = fn(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kmat)
        assert relerr(cpu(fused), cpu(plain)) < 1e-12, fn.__name__
    # ... and the fused pass (hyper-parameter + row-point gradients from one kernel) equals the two separate kernels
    dh_f, dz_f = L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar, epilogue=(2.0, cs, rv, cv), with_points=True)
    assert relerr(cpu(dh_f), cpu(L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kmat))) < 1e-12
    assert relerr(cpu(dz_f), cpu(L.kernel_grad_points(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kmat))) < 1e-12
    dh_p, dz_p = L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar, with_points=True)
    assert relerr(cpu(dz_p), cpu(dz)) < 1e-12 and relerr(cpu(dh_p), cpu(L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar))) < 1e-12
    only_scale = L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, Kbar, epilogue=(0.5, cs, None, None))
    assert relerr(cpu(only_scale), cpu(L.kernel_grad(kind, mode, zd, xd, hd, P, Qk, fz, fx, 0.5 * Kbar * cs[:, None, :]))) < 1e-12
    # two latent GPs per window share one row of points (divA = 2): the wrapper sums their contributions
    h2 = hd[:, :, :].reshape(W * P, 1, -1).contiguous()
    zs = zd.clone().requires_grad_(True)
    K2 = KernelMatrix.apply(h2, zs, xd, kind, mode, 0.0, True)
    Kb3 = torch.randn(W * P, M, N, dtype=DT, device='cuda')
    (K2 * Kb3).sum().backward()
    f2z = L.features(zd, h2, 1, Qk) if kind == 'mercer_m12' else None
    f2x = L.features(xd, h2, 1, Qk) if kind == 'mercer_m12' else None
    each = L.kernel_grad_points(kind, mode, zd, xd, h2, 1, Qk, f2z, f2x, Kb3)
    assert relerr(cpu(zs.grad), cpu(each.view(W, P, M).sum(1))) < 1e-13


@pytest.mark.parametrize('Q,P', [(13, 1), (23, 2), (10, 2), (3, 1)])
def test_builder_and_grad_many_partials(L, Q, P):
    """More partials than one register chunk of the gradient kernel holds (10): every chunk contributes to ALL four
    gradient blocks; need_ef = 0 (fixed energies / frequencies) gives the same variance / lengthscale gradients."""
    rng = np.random.default_rng(Q)
    W, M, N = 2, 45, 333
    x = np.arange(W * N).reshape(W, N) / 16000.
    z = x[:, ::7][:, :M].copy()
    hyp = np.zeros((W, P, 2 + 2 * Q))
    hyp[:, :, 0] = rng.uniform(0.5, 2.0, (W, P)); hyp[:, :, 1] = rng.uniform(0.005, 0.05, (W, P))
    hyp[:, :, 2:2 + Q] = rng.uniform(0.05, 1.0, (W, P, Q)); hyp[:, :, 2 + Q:] = rng.uniform(100, 4000, (W, P, Q))
    zd, xd, hd = dev(z), dev(x), dev(hyp)
    fz, fx = L.features(zd, hd, P, Q), L.features(xd, hd, P, Q)
    K = L.kernel_build('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx)
    Kbar = torch.randn(W, M, N, dtype=DT, device='cuda')
    dh = L.kernel_grad('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx, Kbar)
    dh0 = L.kernel_grad('mercer_m12', 'reference', zd, xd, hd, P, Q, fz, fx, Kbar, need_ef=False)
    assert relerr(cpu(dh0[:, :, :2]), cpu(dh[:, :, :2])) < 1e-12 and float(dh0[:, :, 2:].abs().max()) == 0.0
    with clean_l():
        for w in range(W):
            ht = torch.as_tensor(hyp[w]).clone().requires_grad_(True)
            kerns = [{'kind': 'mercer_m12', 'variance': ht[p, 0], 'lengthscales': ht[p, 1], 'energy': ht[p, 2:2 + Q],
                      'frequency': ht[p, 2 + Q:]} for p in range(P)]
            Kref = KR.K(kerns, torch.as_tensor(z[w]).reshape(-1, 1), torch.as_tensor(x[w]).reshape(-1, 1))
            (Kref * cpu(Kbar[w])).sum().backward()
            assert relerr(cpu(K[w]), Kref.detach()) < 1e-11
            got = cpu(dh[w])
            for c0, c1, nm in ((0, 1, 'var'), (1, 2, 'len'), (2, 2 + Q, 'energy'), (2 + Q, 2 + 2 * Q, 'freq')):
                assert relerr(got[:, c0:c1], ht.grad[:, c0:c1]) < 1e-9, (w, nm)


@pytest.mark.parametrize('t0', [0.0, 10.0, 240.0])
@pytest.mark.parametrize('P,mode', [(1, 'reference'), (1, 'stable'), (6, 'reference')])
def test_lag_histogram_gradient_matches_the_feature_path(L, t0, P, mode):
    """gpx_kernel_grad_lag (inducing points on the sample grid: one pass bins Kbar exp(-r) by integer lag, O(lags x Q)
    tail) against gpx_kernel_grad (per-element Mercer features) on the same inputs, with and without the fused adjoint
    epilogue, at window-local and absolute time stamps.  The two differ only in how the cosine factor's argument is
    rounded (exact lag distance vs. fl(w z) - fl(w x)): <= ulp(t) w_q per element, i.e. ~1e-9 at t = 240 s."""
    from gpitch_b200.batched import grid_lags
    torch.manual_seed(5)
    W, k, N, M, Q, fs = 2, 3, 1000, 80, 5, 16000.0
    batch = W * k
    x = torch.stack([(t0 * fs + w * N + torch.arange(N, dtype=DT)) / fs for w in range(W)]).cuda()
    sel = torch.stack([torch.sort(torch.randperm(N)[:M]).values for _ in range(batch)])          # irregular subsets of the grid
    z = torch.stack([x[r // k][sel[r].cuda()] for r in range(batch)]).contiguous()
    lag = grid_lags(x, z)
    assert lag is not None and lag[0].dtype == torch.int32 and torch.equal(lag[0].cpu().long(), sel)
    hyp = torch.empty(batch, P, 2 + 2 * Q, dtype=DT)
    hyp[:, :, 0] = 0.5 + torch.rand(batch, P)
    hyp[:, :, 1] = 0.01 + 0.05 * torch.rand(batch, P)
    hyp[:, :, 2:2 + Q] = 0.1 + torch.rand(batch, P, Q)
    hyp[:, :, 2 + Q:] = 100.0 + 3000.0 * torch.rand(batch, P, Q)
    hyp = hyp.cuda()
    Kbar = torch.randn(batch, M, N, dtype=DT, device='cuda')
    fz, fx = L.features(z, hyp, P, Q), L.features(x, hyp, P, Q)
    epi = (2.0, torch.randn(batch, N, dtype=DT, device='cuda'), torch.randn(batch, M, dtype=DT, device='cuda'),
           torch.randn(batch, N, dtype=DT, device='cuda'))
    tol = 1e-10 if t0 == 0.0 else (1e-9 if t0 == 10.0 else 2e-8)
    for e in (None, epi):
        ref = L.kernel_grad('mercer_m12', mode, z, x, hyp, P, Q, fz, fx, Kbar, epilogue=e)
        got = L.kernel_grad_lag(mode, z, x, hyp, P, Q, Kbar, lag, epilogue=e)
        for name, cols in (('var', slice(0, 1)), ('len', slice(1, 2)), ('energy', slice(2, 2 + Q)), ('freq', slice(2 + Q, None))):
            assert relerr(cpu(got[:, :, cols]), cpu(ref[:, :, cols])) < tol, (name, e is not None)
    got = L.kernel_grad_lag(mode, z, x, hyp, P, Q, Kbar, lag, need_ef=False)
    assert float(got[:, :, 2:].abs().max()) == 0.0
    # off-grid points / non-uniform grids are detected (the general kernel then runs)
    z_off = z.clone(); z_off[1, 3] += 1e-7
    assert grid_lags(x, z_off) is None
    x_bad = x.clone(); x_bad[0, 10] += 1e-9
    assert grid_lags(x_bad, z) is None


def test_builder_jitter_and_ragged_tile_edges(L):
    rng = np.random.default_rng(1)
    for M in (1, 31, 33, 129):
        z = dev(rng.uniform(0, 0.01, (2, M)))
        hyp = dev(np.tile(np.array([1.2, 0.01, 0.7, 0.3, 200., 410.]), (2, 1, 1)))
        fz = L.features(z, hyp, 1, 2)
        K = L.kernel_build('mercer_m12', 'reference', z, z, hyp, 1, 2, fz, fz, jitter=1e-6)
        kern = KR.make('mercer_m12', 1.2, 0.01, [0.7, 0.3], [200., 410.])
        for w in range(2):
            zt = cpu(z[w]).reshape(-1, 1)
            ref = KR.K(kern, zt) + 1e-6 * torch.eye(M, dtype=DT)
            assert relerr(cpu(K[w]), ref) < 1e-12


# ----------------------------------------------------------------------------------------------- epilogues
def test_scale_rank1_materialises_the_fused_adjoint(L):
    """gpx_scale_rank1 writes out what the gradient kernels' fused epilogue consumes on the fly."""
    b, M, N = 3, 77, 301
    T = torch.randn(b, M, N, dtype=DT, device='cuda')
    cs, cv = torch.randn(b, N, dtype=DT, device='cuda'), torch.randn(b, N, dtype=DT, device='cuda')
    rv = torch.randn(b, M, dtype=DT, device='cuda')
    out = L.scale_rank1(T, cs, rv, cv, alpha=2.0)
    assert relerr(cpu(out), cpu(2.0 * T * cs[:, None, :] + rv[:, :, None] * cv[:, None, :])) < 1e-15


@pytest.mark.parametrize('b,M,N', [(4, 77, 1001), (1, 200, 200), (1, 200, 1600), (80, 77, 1001)])
def test_colstats_rowdot(L, b, M, N):
    """(small launches take the many-CTA variant of the column statistics, large ones the streaming kernel)"""
    torch.manual_seed(3)
    A = torch.randn(b, M, N, dtype=DT, device='cuda'); LT = torch.randn(b, M, N, dtype=DT, device='cuda')
    mu = torch.randn(b, M, dtype=DT, device='cuda'); kd = torch.rand(b, dtype=DT, device='cuda') + 1
    fm, fv = L.cond_colstats(A, LT, mu, kd)
    assert relerr(cpu(fm), cpu(torch.einsum('bmn,bm->bn', A, mu))) < 1e-13
    assert relerr(cpu(fv), cpu(kd[:, None] - (A * A).sum(1) + (LT * LT).sum(1))) < 1e-13
    fm2, fv2 = L.cond_colstats(A, None, mu, kd)
    assert relerr(cpu(fv2), cpu(kd[:, None] - (A * A).sum(1))) < 1e-13
    fm3, fv3 = L.cond_colstats(A, LT, mu, kd, mode=1)
    assert relerr(cpu(fm3), cpu(fm)) < 1e-14 and relerr(cpu(fv3), cpu(kd[:, None] + (A * LT).sum(1))) < 1e-13
    v = torch.randn(b, N, dtype=DT, device='cuda')
    assert relerr(cpu(L.rowdot(A, v)), cpu(torch.einsum('bmn,bn->bm', A, v))) < 1e-13


@pytest.mark.parametrize('P_', [1, 3])
@pytest.mark.parametrize('nl', ['logistic', 'softplus', 'gauss'])
def test_varexp_vs_reference_golden(L, P_, nl):
    g = load_golden('mpdlik_P%d' % P_)
    n = g['Fmu'].shape[0]
    Fmu = dev(g['Fmu'].T.reshape(1, 2 * P_, n)); Fvar = dev(g['Fvar'].T.reshape(1, 2 * P_, n))
    Y = dev(g['Y'].T); noise = dev([float(G.positive_forward(G.positive_backward(torch.tensor(float(g['noise_var']), dtype=DT))))])
    ve, dmu, dvar, dn, pt = L.varexp(Fmu, Fvar, Y, noise, nl, pointwise=True)
    assert relerr(cpu(pt[0]), g['ve_' + nl][:, 0]) < 1e-12
    assert abs(float(ve[0]) - g['ve_' + nl].sum()) < 1e-12 * abs(g['ve_' + nl].sum())
    assert relerr(cpu(dmu[0]).T, g['dFmu_' + nl]) < 1e-11
    assert relerr(cpu(dvar[0]).T, g['dFvar_' + nl]) < 1e-11
    # golden holds d/d(free noise); chain factor of the positive transform: dy/dx = 1 - exp(-(y - 1e-6))... = sigmoid(x)
    xfree = G.positive_backward(torch.tensor(float(g['noise_var']), dtype=DT))
    assert abs(float(dn[0]) * float(torch.sigmoid(xfree)) - float(g['dfree_noise_' + nl][0])) < 1e-10 * abs(float(g['dfree_noise_' + nl][0]))


def test_varexp_batched_vs_oracle(L):
    rng = np.random.default_rng(11)
    P_, W, N = 12, 3, 517
    Fmu = rng.standard_normal((W, 2 * P_, N)) * 2 + 1.0
    Fvar = np.exp(rng.standard_normal((W, 2 * P_, N)) - 1.0)
    Y = rng.standard_normal((W, N)); noise = rng.uniform(0.01, 1.0, W)
    ve, dmu, dvar, dn, _ = L.varexp(dev(Fmu), dev(Fvar), dev(Y), dev(noise), 'logistic')
    for w in range(W):
        v2, dm2, dv2, ds2 = PM.varexp_fwd_bwd(torch.as_tensor(Fmu[w].T.copy()), torch.as_tensor(Fvar[w].T.copy()),
                                              torch.as_tensor(Y[w]), torch.tensor(noise[w], dtype=DT), P_)
        ref = LR.mpdlik_variational_expectations(torch.as_tensor(Fmu[w].T.copy()), torch.as_tensor(Fvar[w].T.copy()),
                                                 torch.as_tensor(Y[w]).reshape(-1, 1), torch.tensor(noise[w], dtype=DT),
                                                 MR.logistic_t, P_).sum()
        assert abs(float(ve[w]) - float(ref)) < 1e-11 * abs(float(ref))
        assert relerr(cpu(dmu[w]).T, dm2) < 1e-11 and relerr(cpu(dvar[w]).T, dv2) < 1e-11
        assert abs(float(dn[w]) - float(ds2)) < 1e-11 * abs(float(ds2))


def test_varexp_88_sources_with_one_dominant_source(L):
    """var_exp contains (sum_i a_i)^2 - sum_i a_i^2 with a_i = E[sigma(g_i)] mu_f_i.  With 88 sources (a full
    piano, gpitch/transcription.py) and one source three orders of magnitude louder than the rest, that difference
    cancels most of S^2; the kernel's value and gradients must still match the oracle's op-for-op evaluation
    (likelihoods.py:56-65) and autograd to 1e-8 -- also in the columns of the quiet sources."""
    rng = np.random.default_rng(88)
    P_, W, N = 88, 2, 300
    Fmu = rng.standard_normal((W, 2 * P_, N)) * 0.5
    Fmu[:, P_ + 17, :] = 400.0 + 50.0 * rng.standard_normal((W, N))        # component mean of the dominant source
    Fmu[:, 17, :] = 6.0                                                      # its activation is fully on
    Fvar = np.exp(rng.standard_normal((W, 2 * P_, N)) - 2.0)
    Y = 400.0 + rng.standard_normal((W, N)); noise = rng.uniform(0.05, 0.5, W)
    ve, dmu, dvar, dn, _ = L.varexp(dev(Fmu), dev(Fvar), dev(Y), dev(noise), 'logistic')
    for w in range(W):
        tm = torch.tensor(Fmu[w].T.copy(), requires_grad=True); tv = torch.tensor(Fvar[w].T.copy(), requires_grad=True)
        tn = torch.tensor(noise[w], dtype=DT, requires_grad=True)
        ref = LR.mpdlik_variational_expectations(tm, tv, torch.as_tensor(Y[w]).reshape(-1, 1), tn, MR.logistic_t, P_).sum()
        ref.backward()
        assert abs(float(ve[w]) - float(ref)) < 1e-10 * abs(float(ref))
        assert relerr(cpu(dmu[w]).T, tm.grad) < 1e-8 and relerr(cpu(dvar[w]).T, tv.grad) < 1e-8
        quiet = [i for i in range(2 * P_) if i not in (17, P_ + 17)]      # the quiet sources' own columns, separately
        assert relerr(cpu(dmu[w]).T[:, quiet], tm.grad[:, quiet]) < 1e-8
        assert relerr(cpu(dvar[w]).T[:, quiet], tv.grad[:, quiet]) < 1e-8
        assert abs(float(dn[w]) - float(tn.grad)) < 1e-8 * abs(float(tn.grad))


def test_gauss_kl_white(L):
    rng = np.random.default_rng(2)
    b, M = 3, 45
    q_mu = rng.standard_normal((b, M)); q_sqrt = np.eye(M) * 0.6 + 0.1 * rng.standard_normal((b, M, M))
    kl, dmu, dLq = L.gauss_kl_white(dev(q_mu), dev(q_sqrt))
    for i in range(b):
        ref = G.gauss_kl(torch.as_tensor(q_mu[i]).reshape(-1, 1), torch.as_tensor(q_sqrt[i])[:, :, None])
        k2, dm2, dl2 = PM.gauss_kl_white(torch.as_tensor(q_mu[i]), torch.tril(torch.as_tensor(q_sqrt[i])))
        assert abs(float(kl[i]) - float(ref)) < 1e-13 * abs(float(ref))
        assert relerr(cpu(dmu[i]), dm2) < 1e-14 and relerr(cpu(dLq[i]), dl2) < 1e-13


@pytest.mark.parametrize('kind', ['mercer_m12', 'matern32'])
@pytest.mark.parametrize('t0,ls', [(0.0, 0.1), (70.0, 0.1), (240.0, 0.01), (10.0, 1.0)])
def test_builder_separable_tiles_vs_oracle(L, kind, t0, ls):
    """Tiles away from the diagonal take the separable exp(-d) u_m v_n path (no per-element sqrt / exp); the result
    must still equal the reference's distance-by-expansion value, also at large absolute time stamps."""
    N, M, Q = 1500, 150, 10
    x = t0 + np.arange(N)[None, :] / 16000.
    z = x[:, ::10][:, :M].copy()
    rng = np.random.default_rng(5)
    e = rng.uniform(0.05, 1.0, Q); f = 261.6 * np.arange(1, Q + 1)
    Qk = 0 if kind == 'matern32' else Q
    hyp = np.concatenate([[1.7, ls], e[:Qk], f[:Qk]])[None, None]
    zd, xd, hd = dev(z), dev(x), dev(hyp)
    fz = L.features(zd, hd, 1, Qk) if Qk else None
    fx = L.features(xd, hd, 1, Qk) if Qk else None
    kern = {'kind': kind, 'variance': torch.tensor(1.7, dtype=DT), 'lengthscales': torch.tensor(ls, dtype=DT),
            'energy': torch.as_tensor(e), 'frequency': torch.as_tensor(f)}
    zt, xt = torch.as_tensor(z[0]).reshape(-1, 1), torch.as_tensor(x[0]).reshape(-1, 1)
    K = L.kernel_build(kind, 'reference', zd, xd, hd, 1, Qk, fz, fx)
    ref = KR.K(kern, zt, xt)
    # 1e-11 of the largest entry, and element-wise relative accuracy wherever the entry is not negligible
    assert relerr(cpu(K[0]), ref) < 1e-11
    big = ref.abs() > 1e-6 * ref.abs().max()
    assert float(((cpu(K[0]) - ref).abs() / ref.abs())[big].max()) < 1e-9
    Kzz = L.kernel_build(kind, 'reference', zd, zd, hd, 1, Qk, fz, fz, jitter=1e-6)
    assert relerr(cpu(Kzz[0]), KR.K(kern, zt) + 1e-6 * torch.eye(M, dtype=DT)) < 1e-11
    Ks = L.kernel_build(kind, 'stable', zd, xd, hd, 1, Qk, fz, fx)
    d = zt - xt.t()
    r = torch.sqrt((d / ls) ** 2 + 1e-12)
    if kind == 'matern32':
        refs = 1.7 * (1 + np.sqrt(3.) * r) * torch.exp(-np.sqrt(3.) * r)
    else:
        refs = 1.7 * torch.exp(-r) * (torch.as_tensor(e)[:, None, None] * torch.cos(
            2 * np.pi * torch.as_tensor(f)[:, None, None] * d[None])).sum(0)
    assert relerr(cpu(Ks[0]), refs) < (1e-9 if t0 > 100 else 1e-10)      # feature phases carry ~1e-10 at t = 240 s


@pytest.mark.parametrize('n,ws', [(1507, 201), (10001, 2001), (2001, 2001), (401, 201)])
def test_overlap_add_device_bit_exact(L, n, ws):
    from gpitch_b200 import window_overlap as WO
    rng = np.random.default_rng(n)
    x = np.arange(n, dtype=np.float64)
    y = rng.standard_normal(n)
    xw, yw = WO.windowed(x, y, ws)
    nm = (ws - 1) // 2 * (len(xw) - 1) + ws
    Y = dev(np.asarray(yw)[:, :, 0])
    assert np.array_equal(cpu(WO.merged_mean_device(Y, ws, nm)).numpy(), WO.merged_mean(yw, ws, nm)[:, 0])
    assert np.array_equal(cpu(WO.merged_variance_device(Y.abs(), ws, nm)).numpy(),
                          WO.merged_variance([np.abs(w) for w in yw], ws, nm)[:, 0])

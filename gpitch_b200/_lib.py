"""ctypes binding of the C-ABI library (include/gpitch_b200.h).  There is NO fallback: if the CUDA library is
missing or no CUDA device is present every compute entry point raises -- the product path never runs on CPU."""
import ctypes as C
import os
import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GPX_LIB', os.path.join(_HERE, 'libgpitch_b200.so'))   # GPX_LIB: kernel experiments only

KIND = {'mercer_m12': 0, 'diff_m12': 1, 'matern32': 2, 'diff_m32': 3}
DIST = {'reference': 0, 'stable': 1}
NLIN = {'logistic': 0, 'softplus': 1, 'gauss': 2}

GEMM_TRANS_A, GEMM_TRANS_B, GEMM_A_LOWER, GEMM_A_UPPER = 1, 2, 4, 8
GEMM_B_LOWER, GEMM_B_UPPER, GEMM_C_LOWER, GEMM_C_MIRROR, GEMM_ZERO_UPPER = 16, 32, 64, 128, 256


class GemmArgs(C.Structure):
    _fields_ = [('A', C.c_void_p), ('B', C.c_void_p), ('C', C.c_void_p),
                ('sA', C.c_longlong), ('sB', C.c_longlong), ('sC', C.c_longlong),
                ('lda', C.c_int), ('ldb', C.c_int), ('ldc', C.c_int),
                ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('batch', C.c_int),
                ('flags', C.c_int),
                ('alpha', C.c_double), ('beta', C.c_double), ('gamma', C.c_double),
                ('alpha_vec', C.c_void_p), ('kweight', C.c_void_p), ('sKw', C.c_longlong),
                ('Aux', C.c_void_p), ('sAux', C.c_longlong), ('ldaux', C.c_int),
                ('colscale', C.c_void_p), ('rowvec', C.c_void_p), ('colvec', C.c_void_p),
                ('sColscale', C.c_longlong), ('sRowvec', C.c_longlong), ('sColvec', C.c_longlong), ('gamma_vec', C.c_void_p)]


EXPORTS = ['gpx_version', 'gpx_set_device', 'gpx_set_hermgauss', 'gpx_feat_rows', 'gpx_features', 'gpx_kernel_build',
           'gpx_kernel_grad', 'gpx_potrf_trinv', 'gpx_gemm', 'gpx_cond_colstats', 'gpx_rowdot', 'gpx_varexp',
           'gpx_gauss_kl_white', 'gpx_launch_count', 'gpx_dmma_peak', 'gpx_scale_rank1', 'gpx_overlap_add', 'gpx_kernel_grad_points', 'gpx_gemm_tma_launch_count', 'gpx_tril_unpack', 'gpx_tril_pack', 'gpx_kernel_grad_lag', 'gpx_potrf_workspace_bytes', 'gpx_kernel_grad_lag_workspace_bytes', 'gpx_kuu_from_kuf', 'gpx_kuu_bar_into_kuf_bar', 'gpx_sgpr_bound', 'gpx_sgpr_bound_workspace_bytes', 'gpx_gauss_kl_white_tril']

_lib = None
_ready_device = None
LAUNCHES = 0          # number of library entry-point calls (each is >= 1 kernel launch); bench.py reports it


def load():
    """dlopen the library (CPU-safe: no CUDA call is made).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('gpitch_b200: %s is missing -- run `python -m gpitch_b200.build` (or '
                               '__graft_entry__.build()); there is no CPU fallback.' % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            getattr(_lib, name).restype = C.c_int
        _lib.gpx_launch_count.restype = C.c_ulonglong
        _lib.gpx_gemm_tma_launch_count.restype = C.c_ulonglong
        _lib.gpx_potrf_workspace_bytes.restype = C.c_longlong
        _lib.gpx_sgpr_bound_workspace_bytes.restype = C.c_longlong
        _lib.gpx_kernel_grad_lag_workspace_bytes.restype = C.c_longlong
    return _lib


def _require_cuda():
    global _ready_device
    if not torch.cuda.is_available():
        raise RuntimeError('gpitch_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    lib = load()
    dev = torch.cuda.current_device()
    if _ready_device != dev:
        _chk(lib.gpx_set_device(C.c_int(dev)), 'gpx_set_device')
        x, w = np.polynomial.hermite.hermgauss(20)           # gpflow.quadrature.hermgauss(20)
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(w / np.sqrt(np.pi), dtype=np.float64)
        _chk(lib.gpx_set_hermgauss(x.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.c_int(20)),
             'gpx_set_hermgauss')
        _ready_device = dev
    return lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed with code %d (%s)' % (what, rc, {-1: 'bad argument', -2: 'CUDA launch failure'}.get(rc, '?')))


def _p(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.dtype in (torch.float64, torch.int32) and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _count():
    global LAUNCHES
    LAUNCHES += 1


def feat_rows(Q):
    return (2 * Q + 3) // 4 * 4


def features(pts, hyp, P, Q):
    """pts [batch/div, n], hyp [batch, P, 2+2Q] -> feat [batch, P, KP, n]."""
    lib = _require_cuda()
    rows, n = pts.shape
    batch = hyp.shape[0]
    assert batch % rows == 0
    div = batch // rows
    feat = torch.empty((batch, P, feat_rows(Q), n), dtype=torch.float64, device=pts.device)
    _chk(lib.gpx_features(_p(pts), C.c_int(n), C.c_int(div), _p(hyp), C.c_int(P), C.c_int(Q), _p(feat), C.c_int(batch),
                          _stream()), 'gpx_features')
    _count()
    return feat


def kernel_build(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, jitter=0.0, out=None):
    lib = _require_cuda()
    rowsA, nA = ptsA.shape
    rowsB, nB = ptsB.shape
    batch = hyp.shape[0]
    assert batch % rowsA == 0 and batch % rowsB == 0
    divA, divB = batch // rowsA, batch // rowsB
    if out is None:
        out = torch.empty((batch, nA, nB), dtype=torch.float64, device=hyp.device)
    ld = out.stride(1)
    with _timed('kernel_build', 8.0 * nA * nB * batch):
      _chk(lib.gpx_kernel_build(C.c_int(KIND[kind]), C.c_int(DIST[mode]), _p(ptsA), C.c_int(nA), C.c_int(divA), _p(ptsB),
                                C.c_int(nB), C.c_int(divB), _p(hyp), C.c_int(P), C.c_int(Q), _p(featA), _p(featB),
                                C.c_void_p(out.data_ptr()), C.c_longlong(out.stride(0)), C.c_int(ld), C.c_double(jitter),
                                C.c_int(batch), _stream()), 'gpx_kernel_build')
    _count()
    return out


def _epi(epilogue):
    """(alpha, colscale [b,nB], rowvec [b,nA] | None, colvec [b,nB] | None) -> ctypes arguments (+ tensors kept alive)."""
    if epilogue is None:
        return (None, None, None, C.c_double(1.0)), ()
    alpha, col, rowv, colv = epilogue
    keep = tuple(None if t is None else t.contiguous() for t in (col, rowv, colv))
    return (_p(keep[0]), _p(keep[1]), _p(keep[2]), C.c_double(alpha)), keep


def kernel_grad(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, Kbar, need_ef=True, epilogue=None, with_points=False):
    """epilogue = (alpha, colscale, rowvec, colvec): consume alpha * colscale[n] * Kbar + rowvec[m] colvec[n] instead.
    with_points: also return the row-point gradient [batch, nA] from the same pass -> (dhyp, dpts)."""
    lib = _require_cuda()
    rowsA, nA = ptsA.shape
    rowsB, nB = ptsB.shape
    batch = hyp.shape[0]
    assert batch % rowsA == 0 and batch % rowsB == 0
    divA, divB = batch // rowsA, batch // rowsB
    assert Kbar.stride(2) == 1
    dhyp = torch.empty((batch, P, 2 + 2 * Q), dtype=torch.float64, device=hyp.device)
    epi_args, _keep = _epi(epilogue)
    dpts = torch.empty((batch, nA), dtype=torch.float64, device=hyp.device) if with_points else None
    with _timed('kernel_grad', 8.0 * nA * nB * batch):
      _chk(lib.gpx_kernel_grad(C.c_int(KIND[kind]), C.c_int(DIST[mode]), _p(ptsA), C.c_int(nA), C.c_int(divA), _p(ptsB),
                             C.c_int(nB), C.c_int(divB), _p(hyp), C.c_int(P), C.c_int(Q), _p(featA), _p(featB),
                             C.c_void_p(Kbar.data_ptr()), C.c_longlong(Kbar.stride(0)), C.c_int(Kbar.stride(1)),
                             _p(dhyp), C.c_int(1 if need_ef else 0), *epi_args, _p(dpts), C.c_int(batch), _stream()),
           'gpx_kernel_grad')
    _count()
    return (dhyp, dpts) if with_points else dhyp


def kernel_grad_lag(mode, ptsA, ptsB, hyp, P, Q, Kbar, lag, need_ef=True, epilogue=None):
    """Hyper-parameter gradient of the MercerMatern12sm cross-covariance for row points on the column grid
    (gpx_kernel_grad_lag).  lag = (iz int32 [rowsA, nA], delta [rowsB], nlag) from batched.grid_lags()."""
    lib = _require_cuda()
    iz, delta, nlag = lag
    rowsA, nA = ptsA.shape
    rowsB, nB = ptsB.shape
    batch = hyp.shape[0]
    assert batch % rowsA == 0 and batch % rowsB == 0 and iz.shape == (rowsA, nA) and delta.shape == (rowsB,)
    assert iz.dtype == torch.int32 and iz.is_contiguous() and delta.is_contiguous() and Kbar.stride(2) == 1
    divA, divB = batch // rowsA, batch // rowsB
    dhyp = torch.empty((batch, P, 2 + 2 * Q), dtype=torch.float64, device=hyp.device)
    work = torch.empty((lib.gpx_kernel_grad_lag_workspace_bytes(C.c_int(nB), C.c_int(P), C.c_int(nlag), C.c_int(batch)) // 8,),
                       dtype=torch.float64, device=hyp.device)
    epi_args, _keep = _epi(epilogue)
    with _timed('kernel_grad', 8.0 * nA * nB * batch):
        _chk(lib.gpx_kernel_grad_lag(C.c_int(DIST[mode]), _p(ptsA), C.c_int(nA), C.c_int(divA), _p(iz), _p(ptsB), C.c_int(nB),
                                     C.c_int(divB), _p(delta), _p(hyp), C.c_int(P), C.c_int(Q),
                                     C.c_void_p(Kbar.data_ptr()), C.c_longlong(Kbar.stride(0)), C.c_int(Kbar.stride(1)),
                                     _p(dhyp), C.c_int(1 if need_ef else 0), *epi_args, _p(work), C.c_int(nlag),
                                     C.c_int(batch), _stream()), 'gpx_kernel_grad_lag')
    _count()
    return dhyp


def kernel_grad_points(kind, mode, ptsA, ptsB, hyp, P, Q, featA, featB, Kbar, epilogue=None):
    """dptsA [batch, nA]: gradient w.r.t. the row points of every batch entry (gpx_kernel_grad_points)."""
    lib = _require_cuda()
    rowsA, nA = ptsA.shape
    rowsB, nB = ptsB.shape
    batch = hyp.shape[0]
    assert batch % rowsA == 0 and batch % rowsB == 0
    divA, divB = batch // rowsA, batch // rowsB
    assert Kbar.stride(2) == 1
    dpts = torch.empty((batch, nA), dtype=torch.float64, device=hyp.device)
    epi_args, _keep = _epi(epilogue)
    with _timed('kernel_grad_points', 8.0 * nA * nB * batch):
      _chk(lib.gpx_kernel_grad_points(C.c_int(KIND[kind]), C.c_int(DIST[mode]), _p(ptsA), C.c_int(nA), C.c_int(divA),
                                    _p(ptsB), C.c_int(nB), C.c_int(divB), _p(hyp), C.c_int(P), C.c_int(Q), _p(featA),
                                    _p(featB), C.c_void_p(Kbar.data_ptr()), C.c_longlong(Kbar.stride(0)),
                                    C.c_int(Kbar.stride(1)), _p(dpts), *epi_args, C.c_int(batch), _stream()),
           'gpx_kernel_grad_points')
    _count()
    return dpts


def potrf_trinv(A):
    """A [batch, M, M] symmetric (overwritten with L).  Returns (L, Linv, info)."""
    lib = _require_cuda()
    batch, M, _ = A.shape
    Linv = torch.empty_like(A)
    work = torch.empty((lib.gpx_potrf_workspace_bytes(C.c_int(M), C.c_int(batch)) // 8,), dtype=torch.float64, device=A.device)
    info = torch.empty((batch,), dtype=torch.int32, device=A.device)
    with _timed('potrf_trinv', 2.0 * batch * (2.0 / 3.0) * M ** 3):          # flops: M^3/3 (potrf) + M^3/3 (inverse)
        _chk(lib.gpx_potrf_trinv(_p(A), C.c_longlong(M * M), C.c_int(M), _p(Linv), C.c_longlong(M * M), C.c_int(M), _p(work),
                                 _p(info), C.c_int(M), C.c_int(batch), _stream()), 'gpx_potrf_trinv')
    _count()
    return A, Linv, info


def _bstride(t):
    return t.stride(0) if t.dim() == 3 else 0


def gemm(A, B, out=None, flags=0, alpha=1.0, beta=0.0, gamma=0.0, alpha_vec=None, kweight=None, aux=None,
         colscale=None, rowvec=None, colvec=None, batch=None, gamma_vec=None):
    """Batched C = op(A) op(B) with the fused epilogue of gpx_gemm.  A, B: [batch, r, c] or [r, c] (shared)."""
    lib = _require_cuda()
    ta, tb = bool(flags & GEMM_TRANS_A), bool(flags & GEMM_TRANS_B)
    if batch is None:
        batch = A.shape[0] if A.dim() == 3 else B.shape[0]
    ar, ac = A.shape[-2], A.shape[-1]
    br, bc = B.shape[-2], B.shape[-1]
    M, K = (ac, ar) if ta else (ar, ac)
    N, K2 = (br, bc) if tb else (bc, br)
    assert K == K2, (A.shape, B.shape, flags)
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    if out is None:
        out = torch.empty((batch, M, N), dtype=torch.float64, device=A.device)
    assert out.stride(-1) == 1
    g = GemmArgs()
    g.A, g.B, g.C = A.data_ptr(), B.data_ptr(), out.data_ptr()
    g.sA, g.sB, g.sC = _bstride(A), _bstride(B), _bstride(out)
    g.lda, g.ldb, g.ldc = A.stride(-2), B.stride(-2), out.stride(-2)
    g.M, g.N, g.K, g.batch, g.flags = M, N, K, batch, flags
    g.alpha, g.beta, g.gamma = alpha, beta, gamma
    keep = [A, B, out]
    for name, t, sname in (('alpha_vec', alpha_vec, None), ('gamma_vec', gamma_vec, None), ('kweight', kweight, 'sKw'), ('colscale', colscale, 'sColscale'),
                           ('rowvec', rowvec, 'sRowvec'), ('colvec', colvec, 'sColvec')):
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float64 and t.stride(-1) == 1
            setattr(g, name, t.data_ptr())
            if sname:
                setattr(g, sname, t.stride(0) if t.dim() == 2 else 0)
            keep.append(t)
    if aux is not None:
        assert aux.stride(-1) == 1
        g.Aux, g.sAux, g.ldaux = aux.data_ptr(), _bstride(aux), aux.stride(-2)
        keep.append(aux)
    with _timed('gemm', gemm_algorithmic_flops(M, N, K, batch, flags) if KernelTimer.active is not None else 0):
        _chk(lib.gpx_gemm(C.byref(g), _stream()), 'gpx_gemm')
    _count()
    return out


def cond_colstats(A, LTA, q_mu, kdiag, mode=0):
    lib = _require_cuda()
    batch, M, N = A.shape
    assert A.is_contiguous() and (LTA is None or (LTA.is_contiguous() and LTA.shape == A.shape))
    fmean = torch.empty((batch, N), dtype=torch.float64, device=A.device)
    fvar = torch.empty_like(fmean)
    with _timed('cond_colstats', 8.0 * M * N * batch * (2 if LTA is not None else 1)):
        _chk(lib.gpx_cond_colstats(_p(A), _p(LTA), C.c_longlong(M * N), C.c_int(N), _p(q_mu), _p(kdiag), _p(fmean), _p(fvar),
                                   C.c_int(M), C.c_int(N), C.c_int(batch), C.c_int(mode), _stream()), 'gpx_cond_colstats')
    _count()
    return fmean, fvar


def scale_rank1(T, colscale, rowvec, colvec, alpha=1.0):
    """alpha * colscale[n] * T + rowvec[m] colvec[n]; T [batch, M, N] contiguous."""
    lib = _require_cuda()
    batch, M, N = T.shape
    assert T.is_contiguous()
    out = torch.empty_like(T)
    _chk(lib.gpx_scale_rank1(_p(T), C.c_longlong(M * N), C.c_int(N), _p(colscale), _p(rowvec), _p(colvec),
                             C.c_double(alpha), _p(out), C.c_int(M), C.c_int(N), C.c_int(batch), _stream()), 'gpx_scale_rank1')
    _count()
    return out


def overlap_add(Y, win, n):
    """Y [W, ws] per-window predictions, win [ws] weights (device) -> merged stream [n]."""
    lib = _require_cuda()
    W, ws = Y.shape
    out = torch.empty((n,), dtype=torch.float64, device=Y.device)
    _chk(lib.gpx_overlap_add(_p(Y.contiguous()), _p(win.contiguous()), C.c_int(W), C.c_int(ws), C.c_int(n), _p(out), _stream()),
         'gpx_overlap_add')
    _count()
    return out


def rowdot(A, v):
    """A [batch, M, N], v [batch, N] or [N] -> [batch, M]."""
    lib = _require_cuda()
    batch, M, N = A.shape
    assert A.is_contiguous() and v.is_contiguous()
    out = torch.empty((batch, M), dtype=torch.float64, device=A.device)
    with _timed('rowdot', 8.0 * M * N * batch):
        _chk(lib.gpx_rowdot(_p(A), C.c_longlong(M * N), C.c_int(N), _p(v), C.c_longlong(N if v.dim() == 2 else 0), _p(out),
                            C.c_int(M), C.c_int(N), C.c_int(batch), _stream()), 'gpx_rowdot')
    _count()
    return out


def varexp(Fmu, Fvar, Y, noise, nlin, need_grad=True, pointwise=False):
    """Fmu, Fvar [W, 2P, N]; Y [W, N]; noise [W].  Returns ve_sum [W], dFmu, dFvar, dnoise, ve_pointwise."""
    lib = _require_cuda()
    W, twoP, N = Fmu.shape
    ve = torch.empty((W,), dtype=torch.float64, device=Fmu.device)
    dFmu = torch.empty_like(Fmu) if need_grad else None
    dFvar = torch.empty_like(Fvar) if need_grad else None
    dn = torch.empty_like(ve) if need_grad else None
    pt = torch.empty((W, N), dtype=torch.float64, device=Fmu.device) if pointwise else None
    with _timed('varexp', 8.0 * N * W * (4 * (twoP // 2) + 1 + (2 * twoP if need_grad else 0))):
      _chk(lib.gpx_varexp(_p(Fmu), _p(Fvar), _p(Y), _p(noise), C.c_int(twoP // 2), C.c_int(W), C.c_int(N), C.c_int(NLIN[nlin]),
                        _p(ve), _p(dFmu), _p(dFvar), _p(dn), _p(pt), _stream()), 'gpx_varexp')
    _count()
    return ve, dFmu, dFvar, dn, pt


def gauss_kl_white(q_mu, q_sqrt, need_grad=True):
    lib = _require_cuda()
    batch, M = q_mu.shape
    kl = torch.empty((batch,), dtype=torch.float64, device=q_mu.device)
    dmu = torch.empty_like(q_mu) if need_grad else None
    dLq = torch.empty_like(q_sqrt) if need_grad else None
    _chk(lib.gpx_gauss_kl_white(_p(q_mu), _p(q_sqrt), C.c_int(M), C.c_int(batch), _p(kl), _p(dmu), _p(dLq), _stream()),
         'gpx_gauss_kl_white')
    _count()
    return kl, dmu, dLq


def kuu_from_kuf(Kuf, iz, pad_diag, jitter):
    """Kuu [batch, M, M] gathered from Kuf [batch, M, N] at the grid columns iz [batch / div, M] of the inducing points."""
    lib = _require_cuda()
    batch, M, N = Kuf.shape
    assert Kuf.stride(2) == 1 and iz.dtype == torch.int32 and iz.is_contiguous() and iz.shape[1] == M and batch % iz.shape[0] == 0
    Kuu = torch.empty((batch, M, M), dtype=torch.float64, device=Kuf.device)
    _chk(lib.gpx_kuu_from_kuf(C.c_void_p(Kuf.data_ptr()), C.c_longlong(Kuf.stride(0)), C.c_int(Kuf.stride(1)), _p(iz),
                              C.c_int(batch // iz.shape[0]), C.c_int(M), _p(pad_diag.contiguous()), C.c_double(jitter), _p(Kuu),
                              C.c_int(batch), _stream()), 'gpx_kuu_from_kuf')
    _count()
    return Kuu


def kuu_bar_into_kuf_bar(Kuu_bar, iz, Kuf_bar):
    """Kuf_bar[:, :, iz_j] += Kuu_bar[:, :, j] in place (adjoint of kuu_from_kuf)."""
    lib = _require_cuda()
    batch, M, N = Kuf_bar.shape
    assert Kuf_bar.stride(2) == 1 and Kuu_bar.is_contiguous() and Kuu_bar.shape == (batch, M, M)
    _chk(lib.gpx_kuu_bar_into_kuf_bar(_p(Kuu_bar), _p(iz), C.c_int(batch // iz.shape[0]), C.c_int(M),
                                      C.c_void_p(Kuf_bar.data_ptr()), C.c_longlong(Kuf_bar.stride(0)), C.c_int(Kuf_bar.stride(1)),
                                      C.c_int(batch), _stream()), 'gpx_kuu_bar_into_kuf_bar')
    _count()
    return Kuf_bar


def sgpr_bound(kind, mode, x, y, z, hyp, noise, jitter=1e-6, reg=False, lag=None, need_grad=True, need_ef=True):
    """gpx_sgpr_bound: collapsed SGPRSS bound (+ gradients) of W windows in one C call -> (bound [W], dhyp, dnoise, info [2, W])."""
    lib = _require_cuda()
    W, N = x.shape
    M = z.shape[1]
    P, Q = hyp.shape[1], (hyp.shape[2] - 2) // 2
    iz, delta, nlag = lag if lag is not None else (None, None, 0)
    nbytes = lib.gpx_sgpr_bound_workspace_bytes(C.c_int(KIND[kind]), C.c_int(N), C.c_int(M), C.c_int(P), C.c_int(Q), C.c_int(W),
                                                C.c_int(1 if need_grad else 0), C.c_int(nlag))
    assert nbytes >= 0
    work = torch.empty((nbytes // 8,), dtype=torch.float64, device=x.device)
    bound = torch.empty((W,), dtype=torch.float64, device=x.device)
    dhyp = torch.empty_like(hyp) if need_grad else None
    dnoise = torch.empty_like(noise) if need_grad else None
    info = torch.empty((2, W), dtype=torch.int32, device=x.device)
    _chk(lib.gpx_sgpr_bound(C.c_int(KIND[kind]), C.c_int(DIST[mode]), _p(x), _p(y), _p(z), C.c_int(N), C.c_int(M), C.c_int(W),
                            _p(hyp), C.c_int(P), C.c_int(Q), _p(noise), C.c_double(jitter), C.c_int(1 if reg else 0),
                            C.c_int(1 if need_ef else 0), _p(iz),
                            _p(delta), C.c_int(nlag), _p(bound), _p(dhyp), _p(dnoise), _p(info), _p(work), _stream()),
         'gpx_sgpr_bound')
    _count()
    return bound, dhyp, dnoise, info


def tril_unpack(packed, M):
    """packed [..., M (M + 1) / 2] lower triangles (row-major) -> dense [..., M, M] with exact zeros above the diagonal."""
    lib = _require_cuda()
    lead = packed.shape[:-1]
    assert packed.shape[-1] == M * (M + 1) // 2 and packed.is_contiguous()
    batch = int(np.prod(lead)) if len(lead) else 1
    dense = torch.empty(tuple(lead) + (M, M), dtype=torch.float64, device=packed.device)
    _chk(lib.gpx_tril_unpack(_p(packed), _p(dense), C.c_int(M), C.c_int(batch), _stream()), 'gpx_tril_unpack')
    _count()
    return dense


def tril_pack(dense):
    """dense [..., M, M] -> packed lower triangles [..., M (M + 1) / 2] (the strict upper triangle is dropped)."""
    lib = _require_cuda()
    M = dense.shape[-1]
    lead = dense.shape[:-2]
    assert dense.shape[-2] == M and dense.is_contiguous()
    batch = int(np.prod(lead)) if len(lead) else 1
    packed = torch.empty(tuple(lead) + (M * (M + 1) // 2,), dtype=torch.float64, device=dense.device)
    _chk(lib.gpx_tril_pack(_p(dense), _p(packed), C.c_int(M), C.c_int(batch), _stream()), 'gpx_tril_pack')
    _count()
    return packed


def gauss_kl_white_tril(q_mu, q_sqrt):
    """(kl [batch], tril(q_sqrt) [batch, M, M]) in one pass over q_sqrt."""
    lib = _require_cuda()
    batch, M = q_mu.shape
    kl = torch.empty((batch,), dtype=torch.float64, device=q_mu.device)
    Lq = torch.empty_like(q_sqrt)
    _chk(lib.gpx_gauss_kl_white_tril(_p(q_mu), _p(q_sqrt), C.c_int(M), C.c_int(batch), _p(kl), _p(Lq), _stream()),
         'gpx_gauss_kl_white_tril')
    _count()
    return kl, Lq


def launch_count():
    """Kernels launched by the library so far in this process."""
    return int(load().gpx_launch_count())


def gemm_tma_launch_count():
    """gpx_gemm launches so far that ran the TMA + mbarrier kernel."""
    return int(load().gpx_gemm_tma_launch_count())


def dmma_peak(reps=5):
    """Measured FP64 tensor-pipe peak (TFLOP/s) of the current device."""
    lib = _require_cuda()
    out = C.c_double(0.0)
    _chk(lib.gpx_dmma_peak(C.c_int(reps), C.byref(out), _stream()), 'gpx_dmma_peak')
    return out.value


class KernelTimer(object):
    """Optional per-launch CUDA-event timing of the library's entry points (bench.py roofline leg).  While active,
    every call records (algorithmic units, start event, stop event) under its entry-point name: flops for
    gpx_gemm, bytes for the HBM-bound kernels (SURVEY.md 8(d) per-unit figures)."""
    active = None

    def __init__(self):
        self.records = {}

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *a):
        KernelTimer.active = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            out[name] = {'units': sum(r[0] for r in recs), 'ms': sum(r[1].elapsed_time(r[2]) for r in recs),
                         'launches': len(recs)}
        return out


class _timed(object):
    def __init__(self, name, units):
        self.tm = KernelTimer.active
        self.name, self.units = name, units

    def __enter__(self):
        if self.tm is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.tm is not None:
            self.e1.record()
            self.tm.records.setdefault(self.name, []).append((self.units, self.e0, self.e1))


def gemm_algorithmic_flops(M, N, K, batch, flags):
    """Algorithmic flop count of one gpx_gemm launch (multiply-add = 2), crediting the triangular structure the
    op has by definition (TRMM M^2 N, SYRK M^2 N, ...), not the padded tiles the kernel executes."""
    tri = bool(flags & (GEMM_A_LOWER | GEMM_A_UPPER | GEMM_B_LOWER | GEMM_B_UPPER))
    both = bool(flags & (GEMM_A_LOWER | GEMM_A_UPPER)) and bool(flags & (GEMM_B_LOWER | GEMM_B_UPPER))
    low = bool(flags & GEMM_C_LOWER)
    f = 1.0
    if tri and low:
        f = 1.0 / 6.0
    elif both:
        f = 1.0 / 3.0
    elif tri or low:
        f = 0.5
    return 2.0 * M * N * K * batch * f

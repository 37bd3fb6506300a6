"""CPU: host-side logic of the product package (window geometry, parameter containers, API surface), the C-ABI
library's exported symbols, the no-fallback rule, and the multi-process sharding helper (gloo, world_size 2)."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def test_window_overlap_bit_exact_vs_reference_golden():
    from gpitch_b200 import window_overlap as WO
    g = load_golden('windows')
    x, y, ws = g['x'], g['y'], int(g['ws'])
    xw, yw = WO.windowed(x, y, ws)
    assert np.array_equal(np.asarray(xw), g['xw']) and np.array_equal(np.asarray(yw), g['yw'])
    n = (ws - 1) // 2 * (len(xw) - 1) + ws
    keep = [w.copy() for w in yw]
    assert np.array_equal(WO.merged_mean(yw, ws, n), g['merged_mean'])
    assert all(np.array_equal(a, b) for a, b in zip(keep, yw))          # inputs are not mutated
    assert np.array_equal(WO.merged_variance([np.abs(w) for w in yw], ws, n), g['merged_variance'])
    assert np.array_equal(WO.merged_x(xw, ws), g['merged_x'])
    xs, ys = WO.segmented(x, y, window_size=300, aug=False)
    assert np.array_equal(np.asarray(xs), g['seg_x']) and np.array_equal(np.asarray(ys), g['seg_y'])
    xa, ya = WO.segmented(x, y, window_size=300, aug=True)
    assert np.array_equal(np.asarray(xa), g['aug_x']) and np.array_equal(np.asarray(ya), g['aug_y'])


def test_window_geometry_table():
    from gpitch_b200 import window_overlap as WO
    for n, ws, nw, start1, start_last, end_last in load_golden('window_geometry')['table']:
        xx = np.arange(n, dtype=np.float64)
        a, _ = WO.windowed(xx, xx, int(ws))
        assert len(a) == nw and int(a[-1][0, 0]) == start_last and int(a[-1][-1, 0]) == end_last


def test_hann_cola_and_edge_cases():
    from gpitch_b200 import window_overlap as WO
    ws = 2001
    x = np.arange(10 * 1000 + 1, dtype=np.float64)
    xw, yw = WO.windowed(x, np.ones_like(x), ws)
    assert len(xw) == 9
    m = WO.merged_mean(yw, ws, x.size)
    assert np.max(np.abs(m - 1.0)) < 1e-15                    # hann(2001) at hop 1000 is COLA
    v = WO.merged_variance(yw, ws, x.size)
    assert abs(v[1500, 0] - 0.5) < 1e-12                      # hann^2 dips to 0.5 mid-overlap
    assert np.array_equal(WO.merged_x(xw, ws)[:, 0], x)
    one, _ = WO.windowed(x[:ws], x[:ws], ws)                  # exactly one window: the reference's if/elif leaves the
    from oracle import window_overlap_ref as WR               # right half Hann-tapered and the centre sample at 0
    assert len(one) == 1 and np.array_equal(WO.merged_mean(one, ws, ws), WR.merged_mean(one, ws, ws))
    assert WO.merged_mean(one, ws, ws)[1000, 0] == 0.0
    assert WO.segmented(x[:100], x[:100], window_size=300)[0] == []


def test_init_liv_and_init_iv_vs_reference_golden():
    from gpitch_b200 import init_models as IM
    g = load_golden('init_models')
    z, yf = IM.init_liv(g['xs'], g['ys'], num_sources=2, win_size=9, thres=0.05, dec=2)
    assert len(z) == 2 and len(z[0]) == 2
    assert np.array_equal(z[0][0], g['liv_z']) and np.array_equal(z[1][1], g['liv_z']) and np.array_equal(yf, g['liv_y'])
    ziv = IM.init_iv(g['xs'], 2, 400, 800, 16000)
    assert np.array_equal(ziv[0][0], g['iv_za']) and np.array_equal(ziv[1][1], g['iv_zc'])
    assert g['wav_liv_z'].shape == (109, 1)               # known answer of demo_modgp-real-audio.ipynb cell 5
    wav = '/root/reference/demos/data/011PFNOF_M60_train.wav'
    if os.path.exists(wav):                                 # build container only: re-derive it from the shipped audio
        from scipy.io import wavfile
        fs, y = wavfile.read(wav)
        y = y.astype(np.float64).reshape(-1, 1)
        x = np.linspace(0., (y.size - 1.) / fs, y.size).reshape(-1, 1)
        zw, yw = IM.init_liv(x, y, win_size=31, thres=0.033, dec=9)
        assert np.array_equal(zw[0][0], g['wav_liv_z']) and np.array_equal(yw, g['wav_liv_y'])
    Z, counts = IM.pad_inducing([np.arange(5.) * 1e-3, np.arange(3.) * 1e-3])
    assert Z.shape == (2, 5) and counts.tolist() == [5, 3] and Z[1, 3] > 999 and Z[1, 4] - Z[1, 3] == 1e3


def test_transforms_match_recalled_gpflow():
    from gpitch_b200.param import transforms
    from oracle import gpflow_ref as G
    y = np.array([1e-3, 0.05, 1.0, 3.5, 261.6255653005986])
    x = transforms.positive.backward(y)
    assert np.array_equal(x, G.positive_backward(torch.as_tensor(y)).numpy())
    assert np.array_equal(transforms.positive.forward(x), G.positive_forward(torch.as_tensor(x)).numpy())
    assert np.allclose(transforms.positive.dforward(x), torch.sigmoid(torch.as_tensor(x)).numpy(), rtol=1e-15)
    lg = transforms.Logistic(0.5, 2.0)
    assert np.allclose(lg.forward(lg.backward(np.array([0.7, 1.9]))), [0.7, 1.9], rtol=1e-14)


def test_api_surface_and_param_tree():
    import gpitch_b200 as gp
    e = [np.array([0.6, 0.4]), np.array([0.7, 0.3])]
    f = [np.array([100., 200.]), np.array([150., 300.])]
    kc = gp.init_kernels.init_kern_com(2, [np.array(0.1), np.array(0.2)], e, f, len_fixed=True)
    ka = gp.init_kernels.init_kern_act(2)
    assert [k.kind for k in kc] == ['mercer_m12'] * 2 and float(ka[0].variance.value) == 3.5
    assert kc[0].lengthscales.fixed and not kc[0].energy[0].fixed          # energies / frequencies stay free
    k2 = gp.Matern12sm(1, variance=1., lengthscales=0.1, energy=e[0], frequency=f[0])
    assert k2.energy[0].fixed and k2.frequency[1].fixed                    # vars_n_freqs_fixed in the constructor
    s = np.sum(kc)
    assert isinstance(s, gp.Add) and len(s.kern_list) == 2 and s.kern_list[1] is kc[1]
    assert np.allclose(s.Kdiag(np.zeros((5, 1))), 2.0)
    m = gp.SGPRSS(np.zeros((10, 1)), np.zeros((10, 1)), s, np.zeros((3, 1)))
    names = [n for n, _ in m.free_params()]
    assert 'kern.kern_list[0].energy[1]' in names and 'likelihood.variance' in names
    assert 'kern.kern_list[0].lengthscales' not in names and 'Z' not in names
    m.likelihood.variance = 0.3
    m.kern.kern_list[0].variance = 2.0
    m.X = np.ones((12, 1))                                                   # DataHolder swap, shape change allowed
    assert float(m.likelihood.variance.value) == 0.3 and m.X.shape == (12, 1)
    x0 = m.get_free_state()
    m.set_state(x0 + 0.1)
    assert np.allclose(m.get_free_state(), x0 + 0.1, rtol=1e-12)
    z = [[np.zeros((4, 1))] * 2, [np.zeros((4, 1))] * 2]
    p = gp.Pdgp(np.zeros((10, 1)), np.zeros((10, 1)), z, [ka, kc])
    assert p.q_sqrt_act[0].shape == (4, 4, 1) and p.num_sources == 2
    assert not p.za[0].fixed and 'za[0]' in [n for n, _ in p.free_params()]     # Params like pdgp.py:80-85 ...
    p.za.fixed = True; p.zc.fixed = True                                          # ... fixed the way demo-modgp.py:40-41 does
    assert p.za[1].fixed and not any(n.startswith('z') for n, _ in p.free_params())
    assert gp.Pdgp(np.zeros((10, 1)), np.zeros((10, 1)), z, [ka, kc], whiten=False).whiten is False


def test_cabi_library_exports_every_declared_symbol():
    from gpitch_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, 'include', 'gpitch_b200.h')).read()
    declared = set(re.findall(r'\b(gpx_[a-z0-9_]+)\s*\(', hdr))
    assert declared and declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gpx_version() >= 100 and lib.gpx_feat_rows(10) == 20 and lib.gpx_feat_rows(5) == 12


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_product_path_has_no_cpu_fallback():
    import gpitch_b200 as gp
    k = gp.MercerMatern12sm(1, energy=np.array([1.]), frequency=np.array([100.]))
    with pytest.raises((RuntimeError, AssertionError)):
        k.K(np.zeros((4, 1)))
    for mod in ('batched', 'functions', '_lib', 'sgpr_ss', 'pdgp', 'kernels', 'likelihoods'):
        src = open(os.path.join(ROOT, 'gpitch_b200', mod + '.py')).read()
        assert 'oracle' not in src, 'product code must never import the oracle (%s)' % mod


def _gloo_worker(rank, world, port, W, out):
    import torch.distributed as dist
    from gpitch_b200.distributed import shard_windows, all_gather_windows
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    lo, hi = shard_windows(W, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
    full = all_gather_windows(local, W)
    if rank == 0:
        torch.save(full, out)
    dist.destroy_process_group()


@pytest.mark.parametrize('W', [7, 8])
def test_window_sharding_all_gather_gloo(tmp_path, W):
    import torch.multiprocessing as mp
    from gpitch_b200.distributed import shard_windows
    assert [shard_windows(7, 2, r) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_windows(3838, 8, r)[1] - shard_windows(3838, 8, r)[0] for r in range(8)] == [480] * 7 + [478]
    out = str(tmp_path / 'g.pt')
    port = 29650 + W
    mp.spawn(_gloo_worker, args=(2, port, W, out), nprocs=2, join=True)
    full = torch.load(out)
    assert full.shape == (W, 3) and torch.equal(full[:, 0], torch.arange(W, dtype=torch.float64))


def test_batched_lbfgs_keeps_per_window_history():
    """The lock-step L-BFGS must cost about what per-window SciPy L-BFGS-B costs, not just reach the same optimum: every
    window keeps its own curvature pairs (a window without a usable pair in some iteration neither inserts a zero pair
    nor loses its older pairs), and a window whose line search hits the rounding floor is frozen instead of spinning."""
    import scipy.optimize as so
    import torch
    from gpitch_b200 import driver
    torch.manual_seed(0)
    W, D = 4, 6
    A = torch.randn(W, D, D, dtype=torch.float64)
    A = A @ A.transpose(1, 2) + 0.5 * torch.eye(D, dtype=torch.float64)
    b = torch.randn(W, D, dtype=torch.float64)

    def val_grad(x):                                   # non-convex toy bounds, one per window (to MAXIMISE)
        x = x.clone().requires_grad_(True)
        q = 0.5 * torch.einsum('wi,wij,wj->w', x, A, x) - (b * x).sum(1) + 0.3 * torch.sin(3 * x).sum(1) + 0.05 * (x ** 4).sum(1)
        (-q).sum().backward()
        return (-q).detach(), {'x': x.grad}
    fs = driver.FreeState({'x': torch.zeros(W, D, dtype=torch.float64)})
    evals = [0]

    def ev(p):
        evals[0] += 1
        return val_grad(p['x'])
    x, hist = driver.lbfgs(fs, ev, torch.zeros(W, D, dtype=torch.float64), 200, gtol=1e-5)
    ref_evals, ref_fun = 0, []
    for w in range(W):
        def fg(xx, w=w):
            v, g = val_grad(torch.as_tensor(xx)[None].repeat(W, 1))
            return -float(v[w]), -g['x'][w].numpy()
        r = so.minimize(fg, np.zeros(D), jac=True, method='L-BFGS-B', options={'gtol': 1e-5, 'ftol': 0})
        ref_evals = max(ref_evals, r.nfev)
        ref_fun.append(r.fun)
    assert np.allclose(hist[-1].numpy(), ref_fun, rtol=1e-8, atol=1e-10)
    assert evals[0] <= 2 * ref_evals, (evals[0], ref_evals)
    evals[0] = 0                                       # unreachable tolerance: stops at the rounding floor, does not spin
    driver.lbfgs(fs, ev, torch.zeros(W, D, dtype=torch.float64), 500, gtol=1e-14)
    assert evals[0] < 400


def _amt_signal(n=9001, fs=16000):
    t = np.arange(n) / float(fs)
    rng = np.random.default_rng(5)
    y = (np.sin(2 * np.pi * 261.6 * t) * np.exp(-((t - 0.15) / 0.08) ** 2) + 0.7 * np.sin(2 * np.pi * 392.0 * t) * np.exp(-((t - 0.4) / 0.1) ** 2)
         + 0.01 * rng.standard_normal(n))
    params = [[np.asarray(0.05), np.asarray(0.05)], [np.array([0.7, 0.3]), np.array([0.6, 0.4])],
              [np.array([261.6, 523.2]), np.array([392.0, 784.0])]]
    return t, y / np.max(np.abs(y)), params


class _StubEngine(object):
    """Stands in for BatchedSGPR on a CPU-only box: the front end's sharding / gathering logic does not care what the
    engine computes, only that per-window results come back in window order."""
    def __init__(self, x, y, z, **kw):
        self.x, self.y, self.z = x, y, z


def _amt_worker(rank, world, port, out):
    import torch.distributed as dist
    import gpitch_b200.transcription as TR
    if world > 1:
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    TR.BatchedSGPR = _StubEngine

    def fake_fit(engine, hyp0, noise0, maxiter=0, train_cols=None, **kw):       # "fitted variance" = a signature of the window
        sig = engine.y.abs().mean(1)[None, :] * torch.arange(1, hyp0.shape[1] + 1, dtype=torch.float64)[:, None]
        return {'hyp': hyp0, 'noise': noise0, 'history': None, 'matrix_var': sig}
    TR.driver.fit_sgpr_windows = fake_fit
    t, y, params = _amt_signal()
    amt = TR.AMT(y, params, pitches=[60, 67], x=t, window_size=2001, overlap=True, device='cpu')
    mv = amt.optimize(maxiter=3)
    if rank == 0:
        np.save(out, mv)
    if world > 1:
        dist.destroy_process_group()


def test_amt_front_end_shards_windows_and_gathers_matrix_var(tmp_path):
    """AMT (array-in): windows by window_overlap.windowed, init_liv inducing points with every 3rd extremum kept
    (transcription.py:229-237), 20 * y fed to the model (transcription.py:255); under torch.distributed every rank fits
    its contiguous block of windows and matrix_var [pitches, windows] is all-gathered -- identical to the 1-process run."""
    import torch.multiprocessing as mp
    import gpitch_b200 as gp
    t, y, params = _amt_signal()
    a = gp.Audio(x=t, y=y, window_size=2001, overlap=True)
    assert len(a.X) == 8 and a.X[0].shape == (2001, 1) and a.wsize == 2001
    one, two = str(tmp_path / 'one.npy'), str(tmp_path / 'two.npy')
    _amt_worker(0, 1, 0, one)
    mp.spawn(_amt_worker, args=(2, 29711, two), nprocs=2, join=True)
    m1, m2 = np.load(one), np.load(two)
    assert m1.shape == (2, 8) and np.array_equal(m1, m2) and np.all(m1[1] == 2 * m1[0]) and np.all(m1 > 0)
    Y = 20.0 * np.stack([w.reshape(-1) for w in a.Y])
    assert np.allclose(m1[0], np.abs(Y).mean(1), rtol=1e-14)

"""GPU parity of the batched engines (SGPR bound / SVGP ELBO, gradients, predictions) against
(1) golden vectors produced by the reference's own source, (2) the oracle on seeded inputs."""
import json
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from oracle import gpflow_ref as G, kernels_ref as KR, methods_ref as MR, pdgp_ref as PR, sgpr_ss_ref as SR

pytestmark = pytest.mark.gpu
DT = torch.float64
T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))


def dev(a):
    return T(a).cuda().contiguous()


def cpu(t):
    return t.detach().cpu()


class clean_l_grad(object):
    """Oracle with the closed-form lengthscale derivative (same forward values) -- see oracle/gpflow_ref.py."""
    def __enter__(self):
        G.CLEAN_LENGTHSCALE_GRAD = True

    def __exit__(self, *a):
        G.CLEAN_LENGTHSCALE_GRAD = False


# d/d(lengthscale) of the reference itself (golden vectors) is only reproducible to about this relative level at
# the golden cases' absolute time stamps (reverse-mode cancellation noise, see oracle/gpflow_ref.py); every other
# parameter block is held to 1e-8, and the lengthscale block is held to 1e-8 against the clean-derivative oracle.
REF_LEN_NOISE = {'t0': 1e-8, 'c1': 1e-6, 't2': 1e-4, 't10': 2e-3, 't240': 1.0}      # (t240: pure noise, see README)


def rt(v):
    """GPflow `positive` transform round trip: value the graph actually sees, and d(constrained)/d(free)."""
    free = G.positive_backward(T(v))
    return G.positive_forward(free).numpy(), torch.sigmoid(free).numpy()


# ------------------------------------------------------------------------------------------ SGPRSS
@pytest.mark.parametrize('tag,reg', [('t0', 0), ('t0', 1), ('t10', 0), ('t10', 1), ('c1', 0), ('t240', 0)])
def test_sgprss_vs_reference_golden(tag, reg):
    """Goldens produced by the reference's own source files (oracle/make_golden.py): small windows at t = 0 / 10 / 240 s
    and the full configs[0] shape (c1: N = 1600, M = 200, P = 3, Q = 10)."""
    from gpitch_b200.batched import BatchedSGPR
    g = load_golden('sgprss_%s_reg%d' % (tag, reg))
    P, Q = g['energy'].shape
    var, dvar = rt(g['variance']); ls, dls = rt(g['lengthscales'])
    e, de = rt(g['energy']); f, df = rt(g['frequency']); nv, dnv = rt(g['noise_var'])
    hyp = np.concatenate([var[:, None], ls[:, None], e, f], 1)[None]            # [1, P, 2+2Q]
    chain = np.concatenate([dvar[:, None], dls[:, None], de, df], 1)[None]
    eng = BatchedSGPR(dev(g['x'].T), dev(g['y'].T), dev(g['z'].T), reg=bool(reg))
    bound, grads = eng.bound(dev(hyp), dev([nv]))
    assert int(eng.last_info.abs().max()) == 0
    assert abs(-float(bound[0]) - float(g['neg_bound'])) < 1e-9 * abs(float(g['neg_bound']))
    gfree = -(cpu(grads['hyp']).numpy() * chain)[0]                                # d(-F)/d(free)
    gn = -float(grads['noise'][0]) * float(dnv)
    names = json.loads(str(g['grad_names']))
    blocks = {}
    for n, ref in zip(names, g['grads']):
        if n == 'likelihood.variance':
            got = gn
        else:
            i = int(n.split('kern_list[')[1].split(']')[0])
            if n.endswith('.variance'):
                got = gfree[i, 0]
            elif n.endswith('.lengthscales'):
                got = gfree[i, 1]
            else:
                q = int(n.rsplit('[', 1)[1][:-1])
                got = gfree[i, 2 + q] if '.energy[' in n else gfree[i, 2 + Q + q]
        tol = REF_LEN_NOISE[tag] if n.endswith('lengthscales') else 1e-8
        blocks.setdefault(n.rsplit('.', 1)[-1].split('[')[0], []).append((got, ref))
        if tag != 't240':          # entry by entry (t = 240 s: per parameter block below -- a 1.5e-3 entry of the frequency
            #                        block differs by 1.6e-10 there with EITHER gradient kernel: the golden's own noise)
            assert abs(got - ref) <= tol * max(abs(ref), np.max(np.abs(g['grads'])) * 1e-6), (n, got, ref)
    for kind, pairs in blocks.items():      # max-norm relative error per parameter block (the north-star criterion)
        got_b, ref_b = np.array([a for a, _ in pairs]), np.array([b for _, b in pairs])
        assert relerr(got_b, ref_b) < (REF_LEN_NOISE[tag] if kind == 'lengthscales' else 1e-8), (kind, relerr(got_b, ref_b))
    # lengthscale block against the clean-derivative oracle (same forward values as the reference)
    with clean_l_grad():
        h = T(hyp[0]).clone().requires_grad_(True)
        kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q],
                  'frequency': h[p, 2 + Q:]} for p in range(P)]
        SR.build_likelihood(T(g['x']), T(g['y']), T(g['z']), kerns, T(nv), reg=bool(reg)).backward()
    assert relerr(cpu(grads['hyp'][0, :, 1]), h.grad[:, 1]) < 1e-8
    mf, vf = eng.predict_f(dev(g['xnew'].T), dev(hyp), dev([nv]))
    assert relerr(cpu(mf[0]), g['predict_f_mean'][:, 0]) < 1e-8 and relerr(cpu(vf[0]), g['predict_f_var'][:, 0]) < 1e-8
    ms, vs = eng.predict_s(dev(g['xnew'].T), dev(hyp), dev([nv]))
    assert relerr(cpu(ms[0]), g['predict_s_mean'][:, :, 0]) < 1e-8 and relerr(cpu(vs[0]), g['predict_s_var'][:, :, 0]) < 1e-8


@pytest.mark.parametrize('use_lag,reg', [(False, False), (True, False), (True, True)])
def test_composite_c_entry_point_equals_the_autograd_path(use_lag, reg):
    """gpx_sgpr_bound -- bound + gradients of W windows in ONE C call, for hosts without torch -- against
    the torch.autograd.Function path (BatchedSGPR.bound), with and without the grid structure of the inducing points, and
    against the oracle."""
    from gpitch_b200 import _lib
    from gpitch_b200.batched import BatchedSGPR, grid_lags
    W, N, M, P, Q = 3, 800, 80, 3, 4
    x, y, z, hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=5, t_origin=False)
    xd, yd, zd, hd, nd = dev(x), dev(y), dev(z), dev(hyp), dev(noise)
    eng = BatchedSGPR(xd, yd, zd, reg=reg)
    eng.lag_grad = 'auto' if use_lag else False
    eng.use_composite = False                      # the torch.autograd.Function path (BatchedSGPR defaults to the C call)
    b_ref, g_ref = eng.bound(hd, nd)
    eng2 = BatchedSGPR(xd, yd, zd, reg=reg)
    assert eng2.use_composite
    b_c, g_c = eng2.bound(hd, nd)                  # ... and the engine's own use of the composite entry point
    assert relerr(cpu(b_c), cpu(b_ref)) < 1e-13 and relerr(cpu(g_c['hyp']), cpu(g_ref['hyp'])) < 1e-11
    b_f, _ = eng2.bound(hd, nd, need_grad=False)
    g_nef = eng2.bound(hd, nd, need_ef=False)[1]['hyp']           # fixed partials: only the Kdiag term reaches the energies
    g_nef_ref = eng.bound(hd, nd, need_ef=False)[1]['hyp']
    assert relerr(cpu(b_f), cpu(b_ref)) < 1e-13 and relerr(cpu(g_nef), cpu(g_nef_ref)) < 1e-11
    assert float(g_nef[:, :, 2 + Q:].abs().max()) == 0.0
    lag = grid_lags(xd, zd) if use_lag else None
    assert (lag is not None) == use_lag
    b, dh, dn, info = _lib.sgpr_bound('mercer_m12', 'reference', xd, yd, zd, hd, nd, jitter=1e-6, reg=reg, lag=lag)
    assert int(info.abs().max()) == 0 and info.shape == (2, W)
    assert relerr(cpu(b), cpu(b_ref)) < 1e-13
    assert relerr(cpu(dh), cpu(g_ref['hyp'])) < 1e-11 and relerr(cpu(dn), cpu(g_ref['noise'])) < 1e-12
    b2, dh2, dn2, _ = _lib.sgpr_bound('mercer_m12', 'reference', xd, yd, zd, hd, nd, jitter=1e-6, reg=reg, need_grad=False)
    assert dh2 is None and relerr(cpu(b2), cpu(b)) < 1e-14
    h = T(hyp[0]).clone().requires_grad_(True); nv = T(noise[0]).clone().requires_grad_(True)
    kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q], 'frequency': h[p, 2 + Q:]}
             for p in range(P)]
    with clean_l_grad():
        ref = SR.build_likelihood(T(x[0]).reshape(-1, 1), T(y[0]).reshape(-1, 1), T(z[0]).reshape(-1, 1), kerns, nv, reg=reg)
        ref.backward()
    assert abs(float(b[0]) - float(ref)) < 1e-8 * abs(float(ref))
    for c0, c1 in ((0, 1), (1, 2), (2, 2 + Q), (2 + Q, 2 + 2 * Q)):
        assert relerr(cpu(dh[0])[:, c0:c1], h.grad[:, c0:c1]) < 1e-8
    assert abs(float(dn[0]) - float(nv.grad)) < 1e-8 * abs(float(nv.grad))


def _rand_sgpr(W, N, M, P, Q, seed, t_origin=True):
    rng = np.random.default_rng(seed)
    x = np.stack([(0.0 if t_origin else w * N / 16000.) + np.arange(N) / 16000. for w in range(W)])
    z = x[:, ::N // M][:, :M].copy()
    y = rng.standard_normal((W, N))
    hyp = np.zeros((W, P, 2 + 2 * Q))
    hyp[:, :, 0] = rng.uniform(0.3, 2.0, (W, P)); hyp[:, :, 1] = rng.uniform(0.01, 0.2, (W, P))
    en = rng.uniform(0.1, 1.0, (W, P, Q)); hyp[:, :, 2:2 + Q] = en / en.sum(-1, keepdims=True)
    f0 = MR.midi2freq(rng.integers(48, 72, (W, P)))
    hyp[:, :, 2 + Q:] = f0[:, :, None] * np.arange(1, Q + 1)
    noise = rng.uniform(0.01, 0.5, W)
    return x, y, z, hyp, noise


@pytest.mark.parametrize('W,N,M,P,Q', [(3, 400, 40, 2, 3), (2, 1600, 200, 3, 10), (2, 801, 67, 1, 4)])
def test_sgpr_bound_and_grad_vs_oracle(W, N, M, P, Q):
    from gpitch_b200.batched import BatchedSGPR
    x, y, z, hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=N)
    eng = BatchedSGPR(dev(x), dev(y), dev(z))
    bound, grads = eng.bound(dev(hyp), dev(noise))
    for w in range(W):
        h = T(hyp[w]).clone().requires_grad_(True); nv = T(noise[w]).clone().requires_grad_(True)
        kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q],
                  'frequency': h[p, 2 + Q:]} for p in range(P)]
        ref = SR.build_likelihood(T(x[w]).reshape(-1, 1), T(y[w]).reshape(-1, 1), T(z[w]).reshape(-1, 1), kerns, nv)
        ref.backward()
        assert abs(float(bound[w]) - float(ref)) < 1e-8 * abs(float(ref)), (w, float(bound[w]), float(ref))
        got = cpu(grads['hyp'][w])
        for c0, c1, nm in ((0, 1, 'var'), (1, 2, 'len'), (2, 2 + Q, 'energy'), (2 + Q, 2 + 2 * Q, 'freq')):
            assert relerr(got[:, c0:c1], h.grad[:, c0:c1]) < 1e-8, (w, nm)
        assert abs(float(grads['noise'][w]) - float(nv.grad)) < 1e-8 * abs(float(nv.grad))


def test_sgpr_full_cov_predictions_vs_oracle():
    """SGPR.build_predict(full_cov=True) and SGPRSS.build_predict_source(full_cov=True) (sgpr_ss.py:99-100): full
    posterior covariances; their diagonals must reproduce the full_cov=False outputs."""
    from gpitch_b200.batched import BatchedSGPR
    W, N, M, P, Q, Ns = 2, 400, 40, 2, 3, 131
    x, y, z, hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=9)
    xnew = x[:, :3 * Ns:3] + 1e-5
    eng = BatchedSGPR(dev(x), dev(y), dev(z))
    mf, cf = eng.predict_f(dev(xnew), dev(hyp), dev(noise), full_cov=True)
    ms, cs = eng.predict_s(dev(xnew), dev(hyp), dev(noise), full_cov=True)
    mf0, vf0 = eng.predict_f(dev(xnew), dev(hyp), dev(noise))
    ms0, vs0 = eng.predict_s(dev(xnew), dev(hyp), dev(noise))
    assert cf.shape == (W, Ns, Ns) and cs.shape == (W, P, Ns, Ns)
    # K(Xnew, Xnew)'s diagonal is var * exp(-sqrt(1e-12)) * sum(e) in the reference (euclid_dist's 1e-12 offset), its
    # Kdiag is var * sum(e): the two variance outputs differ by 1e-6 * Kdiag, in the reference too
    scale = float(cf.abs().max())
    assert float((cf.diagonal(dim1=1, dim2=2) - vf0).abs().max()) < 1e-5 * scale and torch.equal(mf, mf0)
    assert float((cs.diagonal(dim1=2, dim2=3) - vs0).abs().max()) < 1e-5 * scale and torch.equal(ms, ms0)
    for w in range(W):
        h = T(hyp[w]); nv = T(noise[w])
        kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q],
                  'frequency': h[p, 2 + Q:]} for p in range(P)]
        a = [T(v[w]).reshape(-1, 1) for v in (x, y, z, xnew)]
        m_ref, c_ref = SR.predict_f(a[0], a[1], a[2], kerns, nv, a[3], full_cov=True)
        assert relerr(cpu(mf[w]), m_ref[:, 0]) < 1e-8 and relerr(cpu(cf[w]), c_ref[:, :, 0]) < 1e-8
        sm_ref, sc_ref = SR.build_predict_source(a[0], a[1], kerns, nv, a[3], full_cov=True)
        for p in range(P):
            assert relerr(cpu(ms[w, p]), sm_ref[p][:, 0]) < 1e-8 and relerr(cpu(cs[w, p]), sc_ref[p][:, :, 0]) < 1e-8


# ------------------------------------------------------------------------------------------ Pdgp
@pytest.mark.parametrize('P_,late', [(1, False), (2, False), (2, True)])
def test_pdgp_vs_reference_golden(P_, late):
    from gpitch_b200.batched import BatchedPdgp
    g = load_golden('pdgp_P%d_whiten1%s' % (P_, '_t240' if late else ''))
    Q = g['energy'].shape[1]
    va, dva = rt(g['variance_act']); la, dla = rt(g['lengthscales_act'])
    vc, dvc = rt(g['variance_com']); lc, dlc = rt(g['lengthscales_com'])
    e, de = rt(g['energy']); f, df = rt(g['frequency']); nv, dnv = rt(g['noise_var'])
    act_hyp = np.stack([va, la], 1)[None]; act_chain = np.stack([dva, dla], 1)[None]
    com_hyp = np.concatenate([vc[:, None], lc[:, None], e, f], 1)[None]
    com_chain = np.concatenate([dvc[:, None], dlc[:, None], de, df], 1)[None]
    z = np.tile(g['z'].T[None], (1, P_, 1))
    eng = BatchedPdgp(dev(g['x'].T), dev(g['y'].T), dev(z), dev(z))
    args = (dev(act_hyp), dev(com_hyp), dev(g['q_mu_act'][None, :, :, 0]), dev(g['q_sqrt_act'][None, :, :, :, 0]),
            dev(g['q_mu_com'][None, :, :, 0]), dev(g['q_sqrt_com'][None, :, :, :, 0]), dev([nv]))
    elbo, grads = eng.elbo(*args)
    assert int(eng.last_info.abs().max()) == 0
    assert abs(-float(elbo[0]) - float(g['neg_elbo'])) < 1e-9 * abs(float(g['neg_elbo']))
    ga = -(cpu(grads['act_hyp']).numpy() * act_chain)[0]
    gc = -(cpu(grads['com_hyp']).numpy() * com_chain)[0]
    names = json.loads(str(g['grad_names'])); sizes = g['grad_sizes']; off = 0
    for n, s in zip(names, sizes):
        ref = g['grads'][off:off + s]; off += s
        i = int(n.split('[')[1].split(']')[0]) if '[' in n else 0
        if n == 'likelihood.variance':
            got = np.array([-float(grads['noise'][0]) * float(dnv)])
        elif n.startswith('kern_act'):
            got = np.array([ga[i, 0] if n.endswith('variance') else ga[i, 1]])
        elif n.startswith('kern_com'):
            if n.endswith('.variance'):
                got = np.array([gc[i, 0]])
            elif n.endswith('.lengthscales'):
                got = np.array([gc[i, 1]])
            else:
                q = int(n.rsplit('[', 1)[1][:-1])
                got = np.array([gc[i, 2 + q] if '.energy[' in n else gc[i, 2 + Q + q]])
        else:
            key = n.split('[')[0]
            got = -cpu(grads[key][0, i]).numpy().ravel()
        if np.max(np.abs(ref)) == 0:
            assert np.max(np.abs(got)) == 0, n
            continue
        tol = REF_LEN_NOISE['t240' if late else 't2'] if n.endswith('lengthscales') else 1e-8     # x is at t = 2 s / 240 s
        assert relerr(got, ref) < tol, (n, relerr(got, ref))
    ma, va_, mc, vc_, ms = eng.predict(dev(g['xnew'].T), *args[:6])
    for got, key in ((ma, 'mean_act'), (va_, 'var_act'), (mc, 'mean_com'), (vc_, 'var_com'), (ms, 'mean_source')):
        assert relerr(cpu(got[0]), g[key][:, :, 0]) < 1e-8, key


@pytest.mark.parametrize('W,N,M,P,Q', [(3, 300, 30, 2, 3), (2, 1000, 100, 3, 10)])
def test_pdgp_elbo_and_grad_vs_oracle(W, N, M, P, Q):
    from gpitch_b200.batched import BatchedPdgp
    rng = np.random.default_rng(N + P)
    x, y, z, com_hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=N + 1)
    com_hyp[:, :, 1] = rng.uniform(0.01, 0.1, (W, P))
    act_hyp = np.stack([rng.uniform(1.0, 4.0, (W, P)), rng.uniform(0.005, 0.05, (W, P))], -1)
    zz = np.tile(z[:, None, :], (1, P, 1))
    qma = 0.5 * rng.standard_normal((W, P, M)) + 1.0; qmc = 0.3 * rng.standard_normal((W, P, M))
    qsa = np.eye(M) * 0.5 + 0.02 * rng.standard_normal((W, P, M, M))
    qsc = np.eye(M) * 0.7 + 0.02 * rng.standard_normal((W, P, M, M))
    eng = BatchedPdgp(dev(x), dev(y), dev(zz), dev(zz), workspace_gb=0.002 if W == 3 else 24.0)   # forces chunking
    elbo, grads = eng.elbo(dev(act_hyp), dev(com_hyp), dev(qma), dev(qsa), dev(qmc), dev(qsc), dev(noise))
    assert int(eng.last_info.abs().max()) == 0
    for w in range(W):
        ah = T(act_hyp[w]).clone().requires_grad_(True); ch = T(com_hyp[w]).clone().requires_grad_(True)
        nv = T(noise[w]).clone().requires_grad_(True)
        tq = {k: [T(v[w, p]).clone().requires_grad_(True) for p in range(P)] for k, v in
              (('qma', qma[..., None]), ('qmc', qmc[..., None]), ('qsa', qsa[..., None]), ('qsc', qsc[..., None]))}
        ka = [{'kind': 'matern32', 'variance': ah[p, 0], 'lengthscales': ah[p, 1]} for p in range(P)]
        kc = [{'kind': 'mercer_m12', 'variance': ch[p, 0], 'lengthscales': ch[p, 1], 'energy': ch[p, 2:2 + Q],
               'frequency': ch[p, 2 + Q:]} for p in range(P)]
        zs = [T(z[w]).reshape(-1, 1)] * P
        ref = PR.build_likelihood(T(x[w]).reshape(-1, 1), T(y[w]).reshape(-1, 1), zs, zs, ka, kc, tq['qma'], tq['qsa'],
                                  tq['qmc'], tq['qsc'], nv)
        ref.backward()
        assert abs(float(elbo[w]) - float(ref)) < 1e-8 * abs(float(ref)), (w, float(elbo[w]), float(ref))
        assert relerr(cpu(grads['act_hyp'][w]), ah.grad) < 1e-8, ('act_hyp', relerr(cpu(grads['act_hyp'][w]), ah.grad))
        got = cpu(grads['com_hyp'][w])
        for c0, c1, nm in ((0, 1, 'var'), (1, 2, 'len'), (2, 2 + Q, 'energy'), (2 + Q, 2 + 2 * Q, 'freq')):
            assert relerr(got[:, c0:c1], ch.grad[:, c0:c1]) < 1e-8, (w, nm, relerr(got[:, c0:c1], ch.grad[:, c0:c1]))
        assert abs(float(grads['noise'][w]) - float(nv.grad)) < 1e-8 * abs(float(nv.grad))
        for p in range(P):
            assert relerr(cpu(grads['q_mu_act'][w, p]), tq['qma'][p].grad[:, 0]) < 1e-8
            assert relerr(cpu(grads['q_mu_com'][w, p]), tq['qmc'][p].grad[:, 0]) < 1e-8
            assert relerr(cpu(grads['q_sqrt_act'][w, p]), tq['qsa'][p].grad[:, :, 0]) < 1e-8
            assert relerr(cpu(grads['q_sqrt_com'][w, p]), tq['qsc'][p].grad[:, :, 0]) < 1e-8


@pytest.mark.parametrize('whiten', [True, False])
@pytest.mark.parametrize('gform', ['auto', True, False])
def test_pdgp_inducing_input_gradients_vs_oracle(whiten, gform):
    """train_z: d ELBO / d za, d zc (Pdgp.za / zc are trainable Params in the reference, pdgp.py:80-85) for several
    windows with a different inducing set per latent GP, both conditional() formulations, whitened or not."""
    from gpitch_b200.batched import BatchedPdgp
    W, N, M, P, Q = 3, 300, 30, 2, 3
    rng = np.random.default_rng(5)
    x, y, z, com_hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=77)
    com_hyp[:, :, 1] = rng.uniform(0.01, 0.1, (W, P))
    act_hyp = np.stack([rng.uniform(1.0, 4.0, (W, P)), rng.uniform(0.005, 0.05, (W, P))], -1)
    za = np.tile(z[:, None, :], (1, P, 1)) + rng.uniform(-1e-5, 1e-5, (W, P, M))
    zc = np.tile(z[:, None, :], (1, P, 1)) + rng.uniform(-1e-5, 1e-5, (W, P, M))
    qma = 0.5 * rng.standard_normal((W, P, M)) + 1.0; qmc = 0.3 * rng.standard_normal((W, P, M))
    qsa = np.eye(M) * 0.5 + 0.02 * rng.standard_normal((W, P, M, M))
    qsc = np.eye(M) * 0.7 + 0.02 * rng.standard_normal((W, P, M, M))
    eng = BatchedPdgp(dev(x), dev(y), dev(za), dev(zc), workspace_gb=0.002, gform=gform, whiten=whiten, train_z=True)
    elbo, grads = eng.elbo(dev(act_hyp), dev(com_hyp), dev(qma), dev(qsa), dev(qmc), dev(qsc), dev(noise))
    assert int(eng.last_info.abs().max()) == 0 and grads['za'].shape == (W, P, M) and grads['zc'].shape == (W, P, M)
    eng0 = BatchedPdgp(dev(x), dev(y), dev(za), dev(zc), gform=gform, whiten=whiten)
    elbo0, grads0 = eng0.elbo(dev(act_hyp), dev(com_hyp), dev(qma), dev(qsa), dev(qmc), dev(qsc), dev(noise))
    # ('auto' certifies every chunk on its own: three one-window chunks here, one three-window chunk in eng0, so the two
    # engines may pick different -- equally valid -- formulations for a window; otherwise chunking changes nothing)
    assert 'za' not in grads0 and relerr(cpu(elbo), cpu(elbo0)) < (1e-11 if gform == 'auto' else 1e-13)
    assert relerr(cpu(grads['com_hyp']), cpu(grads0['com_hyp'])) < (5e-8 if gform == 'auto' else 1e-12)   # (measured 1e-8, whiten=False)
    for w in range(W):
        ah, ch, nv = T(act_hyp[w]), T(com_hyp[w]), T(noise[w])
        tq = {k: [T(v[w, p]) for p in range(P)] for k, v in
              (('qma', qma[..., None]), ('qmc', qmc[..., None]), ('qsa', qsa[..., None]), ('qsc', qsc[..., None]))}
        ka = [{'kind': 'matern32', 'variance': ah[p, 0], 'lengthscales': ah[p, 1]} for p in range(P)]
        kc = [{'kind': 'mercer_m12', 'variance': ch[p, 0], 'lengthscales': ch[p, 1], 'energy': ch[p, 2:2 + Q],
               'frequency': ch[p, 2 + Q:]} for p in range(P)]
        zas = [T(za[w, p]).reshape(-1, 1).clone().requires_grad_(True) for p in range(P)]
        zcs = [T(zc[w, p]).reshape(-1, 1).clone().requires_grad_(True) for p in range(P)]
        ref = PR.build_likelihood(T(x[w]).reshape(-1, 1), T(y[w]).reshape(-1, 1), zas, zcs, ka, kc, tq['qma'], tq['qsa'],
                                  tq['qmc'], tq['qsc'], nv, whiten=whiten)
        ref.backward()
        assert abs(float(elbo[w]) - float(ref)) < 1e-8 * abs(float(ref))
        # a FORCED G-form on the activation group's closely spaced inducing points (cond(Kmm) ~ 1e7, far outside
        # GFORM_COND_MAX) loses ~6e-17 * cond in its explicit inverse; with whiten=False d/dzc is a difference of
        # terms 1e4 times its size, so that forward error shows up at 1e-5 there (measured; the triangular form,
        # which 'auto' certifies for this group, holds 1e-9).  Forced G-form is therefore only pinned loosely.
        tol = (1e-6 if whiten else 1e-3) if gform is True else 1e-7
        for p in range(P):
            assert relerr(cpu(grads['za'][w, p]), zas[p].grad[:, 0]) < tol, (w, p, 'za')
            assert relerr(cpu(grads['zc'][w, p]), zcs[p].grad[:, 0]) < tol, (w, p, 'zc')


# ------------------------------------------------------------------------------------------ BASELINE full sizes
def _c3_problem(W, P=12, N=4000, M=400, Q=10, act_len=1.0):
    from gpitch_b200 import synthetic
    return synthetic.pdgp_problem(W, N, M, P, Q, act_len=act_len)


def test_c3_full_size_window_vs_oracle():
    """One window of the bench workload (configs[2]: N=4000, M=400, P=12, Q=10, Matern32(l=1, var=3.5) activations
    -- Kmm is jitter-dominated, cond ~ 1e9) against the oracle: ELBO and every gradient block."""
    from gpitch_b200.batched import BatchedPdgp
    P, Q = 12, 10
    pr = _c3_problem(1)
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']))
    names = BatchedPdgp.NAMES
    elbo, grads = eng.elbo(*[dev(pr[k]) for k in names])
    assert int(eng.last_info.abs().max()) == 0
    ah = T(pr['act_hyp'][0]).clone().requires_grad_(True); ch = T(pr['com_hyp'][0]).clone().requires_grad_(True)
    nv = T(pr['noise'][0]).clone().requires_grad_(True)
    tq = {k: [T(pr[k][0, p][..., None]).clone().requires_grad_(True) for p in range(P)]
          for k in ('q_mu_act', 'q_mu_com', 'q_sqrt_act', 'q_sqrt_com')}
    ka = [{'kind': 'matern32', 'variance': ah[p, 0], 'lengthscales': ah[p, 1]} for p in range(P)]
    kc = [{'kind': 'mercer_m12', 'variance': ch[p, 0], 'lengthscales': ch[p, 1], 'energy': ch[p, 2:2 + Q],
           'frequency': ch[p, 2 + Q:]} for p in range(P)]
    zs = [T(pr['za'][0, p]).reshape(-1, 1) for p in range(P)]
    with clean_l_grad():
        ref = PR.build_likelihood(T(pr['x'][0]).reshape(-1, 1), T(pr['y'][0]).reshape(-1, 1), zs, zs, ka, kc,
                                  tq['q_mu_act'], tq['q_sqrt_act'], tq['q_mu_com'], tq['q_sqrt_com'], nv)
        ref.backward()
    rel = abs(float(elbo[0]) - float(ref)) / abs(float(ref))
    errs = {'elbo': rel, 'act_hyp': relerr(cpu(grads['act_hyp'][0]), ah.grad), 'com_var': relerr(cpu(grads['com_hyp'][0, :, 0]), ch.grad[:, 0]),
            'com_len': relerr(cpu(grads['com_hyp'][0, :, 1]), ch.grad[:, 1]),
            'com_energy': relerr(cpu(grads['com_hyp'][0, :, 2:2 + Q]), ch.grad[:, 2:2 + Q]),
            'com_freq': relerr(cpu(grads['com_hyp'][0, :, 2 + Q:]), ch.grad[:, 2 + Q:]),
            'noise': abs(float(grads['noise'][0]) - float(nv.grad)) / abs(float(nv.grad)),
            'q_mu_act': max(relerr(cpu(grads['q_mu_act'][0, p]), tq['q_mu_act'][p].grad[:, 0]) for p in range(P)),
            'q_mu_com': max(relerr(cpu(grads['q_mu_com'][0, p]), tq['q_mu_com'][p].grad[:, 0]) for p in range(P)),
            'q_sqrt_act': max(relerr(cpu(grads['q_sqrt_act'][0, p]), tq['q_sqrt_act'][p].grad[:, :, 0]) for p in range(P)),
            'q_sqrt_com': max(relerr(cpu(grads['q_sqrt_com'][0, p]), tq['q_sqrt_com'][p].grad[:, :, 0]) for p in range(P))}
    print('C3 full-size parity (relative, max-norm per block):', {k: '%.2e' % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if v > 1e-8}
    assert not bad, bad


def test_c3_shape_unwhitened_with_trainable_inducing_inputs_vs_oracle():
    """The reference's DEFAULT Pdgp flavour at the named shape (N=4000, M=400, Q=10; P=2 to keep the oracle quick):
    whiten=False and za / zc trainable.  ELBO, hyper-parameter, variational and inducing-input gradients."""
    from gpitch_b200.batched import BatchedPdgp
    P, Q = 2, 10
    pr = _c3_problem(1, P=P, act_len=0.05)           # resolvable activations: cond(Kmm) ~ 1e5 keeps d/dz well defined
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), whiten=False, train_z=True)
    names = BatchedPdgp.NAMES
    elbo, grads = eng.elbo(*[dev(pr[k]) for k in names])
    assert int(eng.last_info.abs().max()) == 0
    ah = T(pr['act_hyp'][0]).clone().requires_grad_(True); ch = T(pr['com_hyp'][0]).clone().requires_grad_(True)
    nv = T(pr['noise'][0]).clone().requires_grad_(True)
    tq = {k: [T(pr[k][0, p][..., None]).clone().requires_grad_(True) for p in range(P)]
          for k in ('q_mu_act', 'q_mu_com', 'q_sqrt_act', 'q_sqrt_com')}
    ka = [{'kind': 'matern32', 'variance': ah[p, 0], 'lengthscales': ah[p, 1]} for p in range(P)]
    kc = [{'kind': 'mercer_m12', 'variance': ch[p, 0], 'lengthscales': ch[p, 1], 'energy': ch[p, 2:2 + Q],
           'frequency': ch[p, 2 + Q:]} for p in range(P)]
    zas = [T(pr['za'][0, p]).reshape(-1, 1).clone().requires_grad_(True) for p in range(P)]
    zcs = [T(pr['zc'][0, p]).reshape(-1, 1).clone().requires_grad_(True) for p in range(P)]
    ref = PR.build_likelihood(T(pr['x'][0]).reshape(-1, 1), T(pr['y'][0]).reshape(-1, 1), zas, zcs, ka, kc,
                              tq['q_mu_act'], tq['q_sqrt_act'], tq['q_mu_com'], tq['q_sqrt_com'], nv, whiten=False)
    ref.backward()
    errs = {'elbo': abs(float(elbo[0]) - float(ref)) / abs(float(ref)),
            'act_var': relerr(cpu(grads['act_hyp'][0, :, 0]), ah.grad[:, 0]),
            'com_var': relerr(cpu(grads['com_hyp'][0, :, 0]), ch.grad[:, 0]),
            'com_energy': relerr(cpu(grads['com_hyp'][0, :, 2:2 + Q]), ch.grad[:, 2:2 + Q]),
            'com_freq': relerr(cpu(grads['com_hyp'][0, :, 2 + Q:]), ch.grad[:, 2 + Q:]),
            'noise': abs(float(grads['noise'][0]) - float(nv.grad)) / abs(float(nv.grad)),
            'q_mu_act': max(relerr(cpu(grads['q_mu_act'][0, p]), tq['q_mu_act'][p].grad[:, 0]) for p in range(P)),
            'q_mu_com': max(relerr(cpu(grads['q_mu_com'][0, p]), tq['q_mu_com'][p].grad[:, 0]) for p in range(P)),
            'q_sqrt_act': max(relerr(cpu(grads['q_sqrt_act'][0, p]), tq['q_sqrt_act'][p].grad[:, :, 0]) for p in range(P)),
            'q_sqrt_com': max(relerr(cpu(grads['q_sqrt_com'][0, p]), tq['q_sqrt_com'][p].grad[:, :, 0]) for p in range(P)),
            'za': max(relerr(cpu(grads['za'][0, p]), zas[p].grad[:, 0]) for p in range(P)),
            'zc': max(relerr(cpu(grads['zc'][0, p]), zcs[p].grad[:, 0]) for p in range(P))}
    print('C3-shape whiten=False + trainable Z parity:', {k: '%.2e' % v for k, v in errs.items()})
    # the window starts at t = 0 here, so the oracle's own autograd noise (lengthscales, inducing inputs) is small;
    # inducing-input gradients are differences of much larger terms (see the small-size test): 1e-7
    bad = {k: v for k, v in errs.items() if v > (1e-7 if k in ('za', 'zc') else 1e-8)}     # measured: za 8e-9, zc 5e-10
    assert not bad, bad


def test_c3_batch_properties_at_full_size():
    """Size-independent properties on a multi-window batch at the named shape: (i) batched == per-window,
    (ii) chunking does not change results, (iii) permuting windows permutes results (independence)."""
    from gpitch_b200.batched import BatchedPdgp
    W = 3
    pr = _c3_problem(W, P=2)
    names = BatchedPdgp.NAMES
    d = {k: dev(pr[k]) for k in names}
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']))
    e_all, g_all = eng.elbo(*[d[k] for k in names])
    eng1 = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), workspace_gb=0.05)   # 1 window / chunk
    assert eng1.chunk_windows() == 1
    e_ch, g_ch = eng1.elbo(*[d[k] for k in names])
    assert relerr(cpu(e_all), cpu(e_ch)) < 1e-13                               # atomics: summation order may differ
    for k in names:
        # (one-window chunks launch few tiles: their long-K products take the split-K path, whose k-summation order differs
        # from the 3-window launch; the jitter-dominated activation group amplifies that last-bit difference: measured 3e-11)
        assert relerr(cpu(g_all[k]), cpu(g_ch[k])) < 1e-9, k
    perm = [2, 0, 1]
    engp = BatchedPdgp(dev(pr['x'][perm]), dev(pr['y'][perm]), dev(pr['za'][perm]), dev(pr['zc'][perm]))
    e_p, g_p = engp.elbo(*[d[k][perm].contiguous() for k in names])
    assert relerr(cpu(e_p), cpu(e_all[perm])) < 1e-14 and relerr(cpu(g_p['q_mu_com']), cpu(g_all['q_mu_com'][perm])) < 1e-12


def test_host_path_dense_and_packed_triangles_match_device_resident():
    """elbo_host (pinned host parameters in, ELBO + gradients out, chunks pipelined over three streams) equals the
    device-resident evaluation, both with dense q_sqrt buffers and with packed lower triangles (half the PCIe bytes);
    gradients of packed buffers come back packed and equal the lower triangle of the dense gradient."""
    from gpitch_b200 import _lib, synthetic
    from gpitch_b200.batched import BatchedPdgp
    W, N, M, P, Q = 5, 600, 72, 2, 3
    pr = synthetic.pdgp_problem(W, N, M, P, Q, act_len=0.02, com_len=0.05)
    names = BatchedPdgp.NAMES
    d = {k: dev(pr[k]) for k in names}
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), workspace_gb=0.02)
    assert 1 < eng.chunk_windows() < W                                            # several chunks -> the pipeline is exercised
    e_dev, g_dev = eng.elbo(*[d[k] for k in names])
    pin = lambda t: torch.empty(t.shape, dtype=DT, pin_memory=True).copy_(t)
    # dense host buffers
    hp = {k: pin(d[k]) for k in names}
    hg = {k: torch.empty(d[k].shape, dtype=DT, pin_memory=True) for k in names}
    he = torch.empty(W, dtype=DT, pin_memory=True)
    eng.elbo_host(hp, he, hg)
    torch.cuda.synchronize()
    assert relerr(he, cpu(e_dev)) < 1e-13
    for k in names:
        assert relerr(hg[k], cpu(g_dev[k])) < 1e-11, k
    # packed lower triangles
    tri = torch.tril_indices(M, M)
    hp2 = {k: pin(_lib.tril_pack(d[k].contiguous()) if k.startswith('q_sqrt') else d[k]) for k in names}
    assert hp2['q_sqrt_act'].shape == (W, P, M * (M + 1) // 2)
    assert torch.equal(hp2['q_sqrt_com'], cpu(d['q_sqrt_com'])[:, :, tri[0], tri[1]])          # row-major lower triangle
    assert torch.equal(cpu(_lib.tril_unpack(dev(hp2['q_sqrt_com'].numpy()), M)), torch.tril(cpu(d['q_sqrt_com'])))
    hg2 = {k: torch.empty(hp2[k].shape, dtype=DT, pin_memory=True) for k in names}
    he2 = torch.empty(W, dtype=DT, pin_memory=True)
    eng.elbo_host(hp2, he2, hg2)
    torch.cuda.synchronize()
    assert relerr(he2, cpu(e_dev)) < 1e-13
    for k in names:
        ref = cpu(g_dev[k])
        if k.startswith('q_sqrt'):
            assert float(torch.triu(ref, 1).abs().max()) == 0.0                   # nothing is lost by dropping the upper part
            ref = ref[:, :, tri[0], tri[1]]
        assert relerr(hg2[k], ref) < 1e-11, k


def test_gform_selection_and_agreement():
    """conditional() in G-form (2 M^2 N products) vs the triangular form (4): 'auto' certifies the G-form per group
    from the Cholesky factors; where it is selected it agrees with the triangular form far inside the parity budget,
    and jitter-dominated groups (Matern32 activations of the bench workload) stay on the triangular form."""
    from gpitch_b200.batched import BatchedPdgp
    pr = _c3_problem(2, P=2)
    names = BatchedPdgp.NAMES
    d = {k: dev(pr[k]) for k in names}
    res = {}
    for g in ('auto', False, True):
        eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), gform=g)
        res[g] = eng.elbo(*[d[k] for k in names])
        if g == 'auto':
            assert eng.gform == {'act': False, 'com': True}, eng.gform
    e0, g0 = res[False]
    ea, ga = res['auto']
    assert relerr(cpu(ea), cpu(e0)) < 1e-11
    for k in names:
        assert relerr(cpu(ga[k]), cpu(g0[k])) < 1e-9, k
    # forcing the G-form on the ill-conditioned activation group costs accuracy (why 'auto' does not pick it)
    e1, g1 = res[True]
    assert relerr(cpu(e1), cpu(e0)) < 1e-5
    # the stable 3-product form ('ha', the default for groups the G-form is not certified for) against GPflow's own
    # operation order ('tri', 4 products) on the jitter-dominated activation group (cond(Kmm) ~ 1e9)
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), gform=False)
    assert eng.stable_form == 'ha'
    eng.stable_form = 'tri'
    et, gt = eng.elbo(*[d[k] for k in names])
    assert relerr(cpu(e0), cpu(et)) < 1e-12
    for k in names:
        assert relerr(cpu(g0[k]), cpu(gt[k])) < 1e-9, k


def test_gform_certified_per_chunk_and_recertified():
    """Every window chunk carries its own G-form certificate (per-window hyper-parameters differ), and
    recertify_gform() -- what the owner of a captured CUDA graph calls periodically -- reports a changed choice."""
    from gpitch_b200.batched import BatchedPdgp
    pr = _c3_problem(4, P=2)
    names = BatchedPdgp.NAMES
    d = {k: dev(pr[k]) for k in names}
    # long component lengthscales in the LAST window only: its Kmm is ill-conditioned, the other chunks' are not
    d['com_hyp'] = d['com_hyp'].clone()
    d['com_hyp'][3, :, 1] *= 1e4
    eng = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), gform='auto')
    eng.chunk_windows = lambda: 1                     # four chunks of one window
    e, g = eng.elbo(*[d[k] for k in names])
    com = {c: v for (grp, c), v in eng._gform_choice.items() if grp == 'com'}
    assert com == {0: True, 1: True, 2: True, 3: False}, com
    assert eng.gform['com'] is False                  # the reported value is the conjunction
    ref = BatchedPdgp(dev(pr['x']), dev(pr['y']), dev(pr['za']), dev(pr['zc']), gform=False)
    e0, g0 = ref.elbo(*[d[k] for k in names])
    assert relerr(cpu(e), cpu(e0)) < 1e-11
    for k in names:
        assert relerr(cpu(g[k]), cpu(g0[k])) < 1e-9, k
    assert eng.recertify_gform(d['act_hyp'], d['com_hyp']) is False
    d['com_hyp'][0, :, 1] *= 1e4
    assert eng.recertify_gform(d['act_hyp'], d['com_hyp']) is True
    assert eng._gform_choice[('com', 0)] is False


def test_ragged_inducing_sets_by_far_padding():
    """init_liv gives every window its own M (separation.py:243-246): padding with far-away inducing points leaves the
    bound, its gradients and the predictions unchanged."""
    from gpitch_b200.batched import BatchedSGPR
    from gpitch_b200.init_models import pad_inducing
    W, N, P, Q = 2, 600, 2, 3
    x, y, z, hyp, noise = _rand_sgpr(W, N, 40, P, Q, seed=3)
    zl = [z[0, :40], z[1, :29]]
    Zp, counts = pad_inducing(zl)
    eng = BatchedSGPR(dev(x), dev(y), dev(Zp))
    b, g = eng.bound(dev(hyp), dev(noise))
    assert int(eng.last_info.abs().max()) == 0
    for w in range(W):
        e1 = BatchedSGPR(dev(x[w:w + 1]), dev(y[w:w + 1]), dev(zl[w][None]))
        b1, g1 = e1.bound(dev(hyp[w:w + 1]), dev(noise[w:w + 1]))
        assert abs(float(b[w]) - float(b1[0])) < 1e-11 * abs(float(b1[0]))
        # variance / energy gradients of the padded problem contain the pads' (exactly cancelling) Kdiag-free terms
        assert relerr(cpu(g['hyp'][w]), cpu(g1['hyp'][0])) < 1e-9
        m, v = eng.predict_f(dev(x), dev(hyp), dev(noise))
        m1, v1 = e1.predict_f(dev(x[w:w + 1]), dev(hyp[w:w + 1]), dev(noise[w:w + 1]))
        assert relerr(cpu(m[w]), cpu(m1[0])) < 1e-10 and relerr(cpu(v[w]), cpu(v1[0])) < 1e-10


def test_batched_optimiser_drivers():
    """Lock-step window-batched L-BFGS / Adam (replacement of the SoSp / AMT per-window loops): every window's bound
    improves monotonically (L-BFGS) and ends where a per-window SciPy L-BFGS-B run on the single-window model ends."""
    import gpitch_b200 as gp
    from gpitch_b200.batched import BatchedSGPR, BatchedPdgp
    from gpitch_b200 import driver
    W, N, M, P, Q = 3, 400, 40, 2, 3
    x, y, z, hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=11)
    hyp[:, :, 0] = 1.0; hyp[:, :, 1] = 0.05; noise[:] = 1.0          # reset point of separation.py:269-277
    eng = BatchedSGPR(dev(x), dev(y), dev(z))
    res = driver.fit_sgpr_windows(eng, dev(hyp), dev(noise), maxiter=40, method='lbfgs')
    h = cpu(res['history'])
    assert bool((h[1:] <= h[:-1] + 1e-9).all()) and bool((h[-1] < h[0] - 1.0).all())
    assert res['matrix_var'].shape == (P, W)
    for w in range(W):
        kerns = gp.init_kernels.init_kern_com(P, [np.asarray(0.05)] * P, [hyp[w, p, 2:2 + Q] for p in range(P)],
                                              [hyp[w, p, 2 + Q:] for p in range(P)], len_fixed=True)
        m = gp.SGPRSS(x[w].reshape(-1, 1), y[w].reshape(-1, 1), np.sum(kerns), z[w].reshape(-1, 1))
        r = m.optimize(maxiter=200)
        assert abs(float(h[-1, w]) - r.fun) < 2e-2 * abs(r.fun) + 0.5, (w, float(h[-1, w]), r.fun)
    # Adam on a small Pdgp batch
    rng = np.random.default_rng(5)
    Mq = 30
    xq, yq, zq, com_hyp, nz = _rand_sgpr(2, 300, Mq, 2, 3, seed=4)
    act_hyp = np.stack([np.full((2, 2), 3.5), np.full((2, 2), 0.02)], -1)
    zz = np.tile(zq[:, None, :], (1, 2, 1))
    p0 = {'act_hyp': dev(act_hyp), 'com_hyp': dev(com_hyp), 'q_mu_act': dev(np.zeros((2, 2, Mq))),
          'q_mu_com': dev(np.zeros((2, 2, Mq))), 'q_sqrt_act': dev(np.tile(np.eye(Mq), (2, 2, 1, 1))),
          'q_sqrt_com': dev(np.tile(np.eye(Mq), (2, 2, 1, 1))), 'noise': dev(nz)}
    e2 = BatchedPdgp(dev(xq), dev(yq), dev(zz), dev(zz))
    out = driver.fit_pdgp_windows(e2, p0, maxiter=15, lr=0.02)
    hh = cpu(out['history'])
    assert bool((hh[-1] < hh[0]).all()) and bool(torch.isfinite(hh).all())


def test_source_reconstruction_like_sosp_predict_s():
    """driver.predict_sources_merged = SoSp's prediction tail + SoSp.predict_s (separation.py:305-313, 341-379): windowed
    test signal -> per-window source posteriors (chunked) -> Hann overlap-add on the device, against the host port of
    window_overlap applied to per-window oracle predictions."""
    from gpitch_b200 import driver, window_overlap
    from gpitch_b200.batched import BatchedSGPR
    rng = np.random.default_rng(4)
    n, ws, P, Q, M = 1101, 201, 2, 3, 20
    xs = (np.arange(n) / 16000.).reshape(-1, 1)
    ys = np.sin(2 * np.pi * 440 * xs) * np.exp(-xs * 20) + 0.1 * rng.standard_normal((n, 1))
    xw, yw = window_overlap.windowed(xs, ys, ws)
    W = len(xw)
    x = np.stack([a[:, 0] for a in xw]); y = np.stack([a[:, 0] for a in yw])
    z = x[:, ::ws // M][:, :M].copy()
    _, _, _, hyp, noise = _rand_sgpr(W, ws, M, P, Q, seed=12)
    eng = BatchedSGPR(dev(x), dev(y), dev(z), workspace_gb=2e-3)          # forces several window slices
    out = driver.predict_sources_merged(eng, dev(hyp), dev(noise), n)
    ms_ref, vs_ref, mf_ref = [[] for _ in range(P)], [[] for _ in range(P)], []
    for w in range(W):
        h = T(hyp[w]); nv = T(noise[w])
        kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q],
                  'frequency': h[p, 2 + Q:]} for p in range(P)]
        a = [T(v[w]).reshape(-1, 1) for v in (x, y, z)]
        sm, sv = SR.build_predict_source(a[0], a[1], kerns, nv, a[0])
        m_f, _ = SR.predict_f(a[0], a[1], a[2], kerns, nv, a[0])
        mf_ref.append(m_f[:, 0].numpy())
        for p in range(P):
            ms_ref[p].append(sm[p].numpy()); vs_ref[p].append(sv[p].numpy())
    assert relerr(cpu(out['mean_f']), np.concatenate(mf_ref)) < 1e-8
    for p in range(P):
        m_host = window_overlap.merged_mean(ms_ref[p], ws, n)
        v_host = window_overlap.merged_variance(vs_ref[p], ws, n)
        assert out['esource'][p][0].shape == (n, 1)
        assert relerr(cpu(out['esource'][p][0]), m_host) < 1e-8 and relerr(cpu(out['esource'][p][1]), v_host) < 1e-8


@pytest.mark.parametrize('N,Q', [(8192, 4), (32768, 10)])
def test_c5_stress_shape_large_m(N, Q):
    """configs[4]: M = 2048 inducing points (128-row GEMM tiles, 32 Cholesky blocks), one window, SGPR bound + gradients
    against the oracle -- at a reduced N = 8192 and at the full named size N = 32768, Q = 10."""
    from gpitch_b200.batched import BatchedSGPR
    W, M, P = 1, 2048, 1
    x, y, z, hyp, noise = _rand_sgpr(W, N, M, P, Q, seed=21)
    hyp[:, :, 1] = 0.02
    eng = BatchedSGPR(dev(x), dev(y), dev(z))
    bound, grads = eng.bound(dev(hyp), dev(noise))
    assert int(eng.last_info.abs().max()) == 0
    h = T(hyp[0]).clone().requires_grad_(True); nv = T(noise[0]).clone().requires_grad_(True)
    kerns = [{'kind': 'mercer_m12', 'variance': h[0, 0], 'lengthscales': h[0, 1], 'energy': h[0, 2:2 + Q], 'frequency': h[0, 2 + Q:]}]
    with clean_l_grad():
        ref = SR.build_likelihood(T(x[0]).reshape(-1, 1), T(y[0]).reshape(-1, 1), T(z[0]).reshape(-1, 1), kerns, nv)
        ref.backward()
    assert abs(float(bound[0]) - float(ref)) < 1e-8 * abs(float(ref))
    got = cpu(grads['hyp'][0])
    for c0, c1, nm in ((0, 1, 'var'), (1, 2, 'len'), (2, 2 + Q, 'energy'), (2 + Q, 2 + 2 * Q, 'freq')):
        assert relerr(got[:, c0:c1], h.grad[:, c0:c1]) < 1e-8, (nm, relerr(got[:, c0:c1], h.grad[:, c0:c1]))
    assert abs(float(grads['noise'][0]) - float(nv.grad)) < 1e-8 * abs(float(nv.grad))


def test_empty_and_single_sample_edges():
    from gpitch_b200.batched import BatchedSGPR
    x, y, z, hyp, noise = _rand_sgpr(1, 64, 8, 1, 2, seed=1)
    eng = BatchedSGPR(dev(x[:0]), dev(y[:0]), dev(z[:0]))                  # no windows at all
    b, g = eng.bound(dev(hyp[:0]), dev(noise[:0]))
    assert b.shape == (0,) and g['hyp'].shape[0] == 0
    eng1 = BatchedSGPR(dev(x[:, :1]), dev(y[:, :1]), dev(z[:, :1]))         # one sample, one inducing point
    b1, _ = eng1.bound(dev(hyp), dev(noise))
    h = T(hyp[0]); kern = [{'kind': 'mercer_m12', 'variance': h[0, 0], 'lengthscales': h[0, 1], 'energy': h[0, 2:4], 'frequency': h[0, 4:]}]
    ref = SR.build_likelihood(T(x[0, :1]).reshape(-1, 1), T(y[0, :1]).reshape(-1, 1), T(z[0, :1]).reshape(-1, 1), kern, T(noise[0]))
    assert abs(float(b1[0]) - float(ref)) < 1e-10 * abs(float(ref))


@pytest.mark.parametrize('Q,M,t_start', [(3, 100, 30.0), (10, 200, 200.0)])
def test_c4_flavour_88_pitch_kernels_on_overlapping_windows(Q, M, t_start):
    """configs[3]: 50 %-overlap windows of ws = 2001 samples (odd leading dimensions -> unaligned copy paths) cut by
    window_overlap.windowed at absolute time, SGPRSS with the Add of 88 pitch kernels (MIDI 21..108) -- a reduced case
    (Q = 3, M = 100, half a minute into the track) and the named shape (Q = 10, M = 200, 200 s into a 4-minute track)."""
    from gpitch_b200.batched import BatchedSGPR
    from gpitch_b200 import window_overlap as WO, synthetic
    P, ws = 88, 2001
    n = 4001
    t = np.arange(n) / 16000. + t_start
    rng = np.random.default_rng(8)
    ysig = rng.standard_normal(n) * 0.1 + np.sin(2 * np.pi * 440 * t)
    xw, yw = WO.windowed(t, ysig, ws)
    W = len(xw)
    assert W == 3
    x = np.stack([a[:, 0] for a in xw]); y = np.stack([a[:, 0] for a in yw])
    z = x[:, ::ws // M][:, :M].copy()
    e, f = synthetic.harmonic_params(list(range(21, 109)), Q)
    hyp = np.tile(np.concatenate([rng.uniform(0.5, 1.5, (P, 1)), 0.05 * np.ones((P, 1)), e, f], 1)[None], (W, 1, 1))
    noise = np.full(W, 0.3)
    eng = BatchedSGPR(dev(x), dev(y), dev(z))
    bound, grads = eng.bound(dev(hyp), dev(noise))
    assert int(eng.last_info.abs().max()) == 0
    for w in (0, W - 1):
        h = T(hyp[w]).clone().requires_grad_(True); nv = T(noise[w]).clone().requires_grad_(True)
        kerns = [{'kind': 'mercer_m12', 'variance': h[p, 0], 'lengthscales': h[p, 1], 'energy': h[p, 2:2 + Q],
                  'frequency': h[p, 2 + Q:]} for p in range(P)]
        with clean_l_grad():
            ref = SR.build_likelihood(T(x[w]).reshape(-1, 1), T(y[w]).reshape(-1, 1), T(z[w]).reshape(-1, 1), kerns, nv)
            ref.backward()
        assert abs(float(bound[w]) - float(ref)) < 1e-8 * abs(float(ref))
        got = cpu(grads['hyp'][w])
        for c0, c1, nm in ((0, 1, 'var'), (1, 2, 'len'), (2, 2 + Q, 'energy'), (2 + Q, 2 + 2 * Q, 'freq')):
            assert relerr(got[:, c0:c1], h.grad[:, c0:c1]) < 1e-8, (w, nm, relerr(got[:, c0:c1], h.grad[:, c0:c1]))
    # predictions, merged on the device, equal the host merge of the per-window predictions bit for bit
    m, v = eng.predict_f(dev(x), dev(hyp), dev(noise))
    nm_ = (ws - 1) // 2 * (W - 1) + ws
    mm = WO.merged_mean_device(m, ws, nm_)
    assert np.array_equal(cpu(mm).numpy(), WO.merged_mean([cpu(m[w]).numpy().reshape(-1, 1) for w in range(W)], ws, nm_)[:, 0])

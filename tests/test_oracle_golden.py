"""CPU: the oracle restatement vs the golden vectors produced by executing the reference's own
source files (oracle/make_golden.py).  Pins every on-disk formula of the hot path."""
import json
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from oracle import gpflow_ref as G, kernels_ref as KR, likelihoods_ref as LR, methods_ref as MR
from oracle import sgpr_ss_ref as S, pdgp_ref as P, window_overlap_ref as WO

T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
TOL = 1e-13


@pytest.mark.parametrize('tag', ['t0', 't10', 't240'])
def test_kernels(tag):
    g = load_golden('kernels_' + tag)
    x, z = T(g['x']), T(g['z'])
    # hyper-parameters reach the graph through GPflow's `positive` transform round trip; at t >= 10 s
    # the reference's distance-by-expansion amplifies that last-ulp change of l to ~1e-9 in K
    rt = lambda v: G.positive_forward(G.positive_backward(T(v)))
    km = KR.make('mercer_m12', rt(g['variance']), rt(g['lengthscales']), rt(g['energy']), rt(g['frequency']))
    kd = KR.make('diff_m12', rt(g['variance']), rt(g['lengthscales']), rt(g['energy']), rt(g['frequency']))
    assert relerr(KR.K(km, z, x), g['mercer_Kzx']) < TOL
    assert relerr(KR.K(km, z), g['mercer_Kzz']) < TOL
    assert relerr(KR.Kdiag(km, x), g['mercer_Kdiag']) < TOL
    assert relerr(KR.phi_features(km, x), g['mercer_phi']) < TOL
    assert relerr(KR.K(kd, z, x), g['diff_Kzx']) < TOL
    assert relerr(KR.K(kd, z), g['diff_Kzz']) < TOL
    assert relerr(KR.Kdiag(kd, x), g['diff_Kdiag']) < TOL


def test_legacy_matern32sm_kernels():
    """Matern32sm / Matern32sml (gpitch/kernels.py:204-318) against vectors from the reference's own source."""
    g = load_golden('legacy_kernels')
    x, z = T(g['x']), T(g['z'])
    ksm = {'kind': 'diff_m32', 'variance': T(1.0), 'lengthscales': T(g['sm_lengthscales']), 'energy': T(g['sm_variance']),
           'frequency': T(g['sm_frequency'])}
    ksml = {'kind': 'diff_m32', 'variance': T(1.0), 'lengthscales': T(g['sml_lengthscales']), 'energy': T(g['sml_variance']),
            'frequency': T(g['sml_frequency'])}
    for tag, k in (('sm', ksm), ('sml', ksml)):
        assert relerr(KR.K(k, z, x), g[tag + '_Kzx']) < TOL
        assert relerr(KR.K(k, z), g[tag + '_Kzz']) < TOL
        assert relerr(KR.Kdiag(k, x), g[tag + '_Kdiag']) < TOL


def test_kernelfit_profile_functions():
    """oracle/kernelfit_ref.py against vectors from the reference's own gpitch/kernelfit.py."""
    from oracle import kernelfit_ref as KF
    g = load_golden('kernelfit')
    assert relerr(KF.approximate_kernel(g['p'], g['x']), g['approx']) < TOL
    assert abs(KF.loss_func(g['p'], g['x'], g['y']) - float(g['loss'])) < TOL * float(g['loss'])
    assert relerr(KF.func(g['x'].reshape(-1), *g['gabor_p']), g['gabor_sum']) < TOL


def test_mercer_equals_difference_form_at_origin():
    """the two kernel classes agree up to their 1e-12 offsets when t is small."""
    g = load_golden('kernels_t0')
    assert relerr(g['mercer_Kzx'], g['diff_Kzx']) < 2e-6      # exp(-1e-6) on coincident points
    off = np.abs(g['x'].T - g['z']) > 1e-9
    assert np.max(np.abs(g['mercer_Kzx'] - g['diff_Kzx'])[off]) < 1e-8


def test_nonlinearities_and_known_answers():
    g = load_golden('nonlin')
    x = g['x']
    assert relerr(MR.logistic(x), g['logistic']) < TOL
    assert relerr(MR.softplus(x), g['softplus']) < TOL
    assert relerr(MR.gaussfun(x), g['gaussfun']) < TOL
    assert relerr(MR.logistic_t(T(x)), g['logistic_tf']) < TOL
    assert relerr(MR.softplus_t(T(x)), g['softplus_tf']) < TOL
    assert relerr(MR.gaussfun_t(T(x)), g['gaussfun_tf']) < TOL
    # demo_modgp-real-audio.ipynb cell 4 known answer
    assert float(g['f0_M60'][0]) == 261.6255653005986 == MR.midi2freq(60)


@pytest.mark.parametrize('P_', [1, 3])
@pytest.mark.parametrize('nl', ['logistic', 'softplus', 'gauss'])
def test_mpdlik(P_, nl):
    g = load_golden('mpdlik_P%d' % P_)
    Fmu = T(g['Fmu']).requires_grad_(True)
    Fvar = T(g['Fvar']).requires_grad_(True)
    free = G.positive_backward(T(g['noise_var'])).reshape(1).requires_grad_(True)
    ve = LR.mpdlik_variational_expectations(Fmu, Fvar, T(g['Y']), G.positive_forward(free), MR.NLIN[nl], P_)
    assert relerr(ve.detach(), g['ve_' + nl]) < TOL
    gr = torch.autograd.grad(ve.sum(), [Fmu, Fvar, free])
    assert relerr(gr[0], g['dFmu_' + nl]) < 1e-12
    assert relerr(gr[1], g['dFvar_' + nl]) < 1e-12
    assert relerr(gr[2], g['dfree_noise_' + nl]) < 1e-12
    lp = LR.mpdlik_logp(T(g['F_' + nl]), T(g['Y']), T(g['noise_var']), MR.NLIN[nl], P_)
    assert relerr(lp, g['logp_' + nl]) < TOL


def test_modlik():
    g = load_golden('modlik')
    ve = LR.modlik_variational_expectations(T(g['Fmu']), T(g['Fvar']), T(g['Y']), T(g['noise_var']), MR.logistic_t)
    assert relerr(ve, g['ve']) < TOL


def _free(v):
    return G.positive_backward(T(v)).clone().requires_grad_(True)


@pytest.mark.parametrize('tag,reg', [('t0', 0), ('t0', 1), ('t10', 0), ('t10', 1), ('c1', 0), ('t240', 0)])
def test_sgprss(tag, reg):
    g = load_golden('sgprss_%s_reg%d' % (tag, reg))
    x, y, z, xnew = T(g['x']), T(g['y']), T(g['z']), T(g['xnew'])
    Pn = g['variance'].shape[0]
    fr = {'v': [_free(g['variance'][i]) for i in range(Pn)], 'l': [_free(g['lengthscales'][i]) for i in range(Pn)],
          'e': [_free(g['energy'][i]) for i in range(Pn)], 'f': [_free(g['frequency'][i]) for i in range(Pn)],
          'n': _free(g['noise_var'])}
    kerns = [{'kind': 'mercer_m12', 'variance': G.positive_forward(fr['v'][i]),
              'lengthscales': G.positive_forward(fr['l'][i]), 'energy': G.positive_forward(fr['e'][i]),
              'frequency': G.positive_forward(fr['f'][i])} for i in range(Pn)]
    nv = G.positive_forward(fr['n'])
    bound = S.build_likelihood(x, y, z, kerns, nv, reg=bool(reg))
    assert abs(float(-bound) - float(g['neg_bound'])) < 1e-12 * abs(float(g['neg_bound']))
    (-bound).backward()
    names = json.loads(str(g['grad_names']))
    got = []
    for n in names:
        if n == 'likelihood.variance':
            got.append(fr['n'].grad.reshape(-1))
            continue
        i = int(n.split('kern_list[')[1].split(']')[0])
        if n.endswith('.variance'):
            got.append(fr['v'][i].grad.reshape(-1))
        elif n.endswith('.lengthscales'):
            got.append(fr['l'][i].grad.reshape(-1))
        else:
            q = int(n.rsplit('[', 1)[1][:-1])
            got.append((fr['e'] if '.energy[' in n else fr['f'])[i].grad[q].reshape(-1))
    assert relerr(torch.cat(got), g['grads']) < 1e-11
    with torch.no_grad():
        mf, vf = S.predict_f(x, y, z, kerns, nv, xnew)
        ms, vs = S.build_predict_source(x, y, kerns, nv, xnew)
    assert relerr(mf, g['predict_f_mean']) < 1e-12 and relerr(vf, g['predict_f_var']) < 1e-12
    assert relerr(torch.stack(ms), g['predict_s_mean']) < 1e-11 and relerr(torch.stack(vs), g['predict_s_var']) < 1e-11


@pytest.mark.parametrize('P_,whiten,zfree', [(1, 1, 0), (1, 0, 0), (2, 1, 0), (2, 0, 0), (2, 1, 1), (2, 0, 1), (2, 1, 't240')])
def test_pdgp(P_, whiten, zfree):
    """zfree: the inducing inputs za / zc stay trainable Params (the reference default, pdgp.py:80-85) and the golden
    carries the reference's autodiff gradients w.r.t. them."""
    late = zfree == 't240'          # whitened, fixed inducing inputs, absolute time stamps at the end of a 4-minute track
    zfree = 0 if late else zfree
    g = load_golden('pdgp_P%d_whiten%d%s' % (P_, whiten, '_t240' if late else ('_zfree' if zfree else '')))
    x, y, z, xnew = T(g['x']), T(g['y']), T(g['z']), T(g['xnew'])
    fr = {'va': [_free(v) for v in g['variance_act']], 'la': [_free(v) for v in g['lengthscales_act']],
          'vc': [_free(v) for v in g['variance_com']], 'lc': [_free(v) for v in g['lengthscales_com']],
          'e': [_free(v) for v in g['energy']], 'f': [_free(v) for v in g['frequency']], 'n': _free(g['noise_var']),
          'qma': [T(v).clone().requires_grad_(True) for v in g['q_mu_act']],
          'qmc': [T(v).clone().requires_grad_(True) for v in g['q_mu_com']],
          'qsa': [T(v).clone().requires_grad_(True) for v in g['q_sqrt_act']],
          'qsc': [T(v).clone().requires_grad_(True) for v in g['q_sqrt_com']]}
    pf = G.positive_forward
    ka = [{'kind': 'matern32', 'variance': pf(fr['va'][i]), 'lengthscales': pf(fr['la'][i])} for i in range(P_)]
    kc = [{'kind': 'mercer_m12', 'variance': pf(fr['vc'][i]), 'lengthscales': pf(fr['lc'][i]),
           'energy': pf(fr['e'][i]), 'frequency': pf(fr['f'][i])} for i in range(P_)]
    zs = [z] * P_
    zas = [z.clone().requires_grad_(bool(zfree)) for _ in range(P_)]
    zcs = [z.clone().requires_grad_(bool(zfree)) for _ in range(P_)]
    kl = P.build_prior_kl(zs, zs, ka, kc, fr['qma'], fr['qsa'], fr['qmc'], fr['qsc'], whiten=bool(whiten))
    assert abs(float(kl) - float(g['prior_kl'])) < 1e-12 * abs(float(g['prior_kl']))
    elbo = P.build_likelihood(x, y, zas, zcs, ka, kc, fr['qma'], fr['qsa'], fr['qmc'], fr['qsc'], pf(fr['n']),
                              whiten=bool(whiten))
    assert abs(float(-elbo) - float(g['neg_elbo'])) < 1e-11 * abs(float(g['neg_elbo']))
    (-elbo).backward()
    names = json.loads(str(g['grad_names']))
    got = []
    for n in names:
        if n == 'likelihood.variance':
            got.append(fr['n'].grad.reshape(-1)); continue
        i = int(n.split('[')[1].split(']')[0])
        if n.startswith('za[') or n.startswith('zc['):
            got.append((zas if n.startswith('za') else zcs)[i].grad.reshape(-1)); continue
        if n.startswith('kern_act'):
            got.append((fr['va'] if n.endswith('variance') else fr['la'])[i].grad.reshape(-1))
        elif n.startswith('kern_com'):
            if n.endswith('.variance'):
                got.append(fr['vc'][i].grad.reshape(-1))
            elif n.endswith('.lengthscales'):
                got.append(fr['lc'][i].grad.reshape(-1))
            else:
                q = int(n.rsplit('[', 1)[1][:-1])
                got.append((fr['e'] if '.energy[' in n else fr['f'])[i].grad[q].reshape(-1))
        else:
            key = {'q_mu_act': 'qma', 'q_mu_com': 'qmc', 'q_sqrt_act': 'qsa', 'q_sqrt_com': 'qsc'}[n.split('[')[0]]
            got.append(fr[key][i].grad.reshape(-1))
    got = torch.cat(got).numpy()
    sizes = g['grad_sizes']
    off = 0
    for n, s in zip(names, sizes):      # per parameter block, max-norm relative
        blk_ref = g['grads'][off:off + s]
        # lengthscale gradients flow through the reference's distance-by-expansion, whose autograd sums three
        # terms of size x~^2/l that cancel to d~^2/l: the reference's OWN value carries ~eps*x~^2/d~^2 noise
        # (1e-8..1e-7 here), so two op-for-op runs that differ in one GEMM summation order disagree at that level.
        # At t = 240 s the same noise is ~3e-4 of the value: the golden (reference source under the shim) and this
        # restatement, both torch fp64 on the same host, already disagree at that level -- nobody can match it to 1e-8.
        tol = (5e-3 if late else 1e-6) if n.endswith('lengthscales') else 1e-9
        if np.max(np.abs(blk_ref)) > 0:
            assert relerr(got[off:off + s], blk_ref) < tol, (n, relerr(got[off:off + s], blk_ref))
        off += s
    with torch.no_grad():
        ma, va, mc, vc, msrc = P.predict_act_n_com(xnew, zs, zs, ka, kc, fr['qma'], fr['qsa'], fr['qmc'], fr['qsc'],
                                                   whiten=bool(whiten))
    for got_l, key in ((ma, 'mean_act'), (va, 'var_act'), (mc, 'mean_com'), (vc, 'var_com'), (msrc, 'mean_source')):
        assert relerr(torch.stack(got_l), g[key]) < 1e-10, key


def test_windows_bit_exact():
    g = load_golden('windows')
    x, y, ws = g['x'], g['y'], int(g['ws'])
    xw, yw = WO.windowed(x, y, ws)
    assert np.array_equal(np.asarray(xw), g['xw']) and np.array_equal(np.asarray(yw), g['yw'])
    n = (ws - 1) // 2 * (len(xw) - 1) + ws
    assert np.array_equal(WO.merged_mean(yw, ws, n), g['merged_mean'])
    assert np.array_equal(WO.merged_variance([np.abs(w) for w in yw], ws, n), g['merged_variance'])
    assert np.array_equal(WO.merged_x(xw, ws), g['merged_x'])
    xs, ys = WO.segmented(x, y, window_size=300, aug=False)
    assert np.array_equal(np.asarray(xs), g['seg_x']) and np.array_equal(np.asarray(ys), g['seg_y'])
    xa, ya = WO.segmented(x, y, window_size=300, aug=True)
    assert np.array_equal(np.asarray(xa), g['aug_x']) and np.array_equal(np.asarray(ya), g['aug_y'])


def test_window_geometry_python2_division():
    for n, ws, nw, start1, start_last, end_last in load_golden('window_geometry')['table']:
        xx = np.arange(n, dtype=np.float64)
        a, _ = WO.windowed(xx, xx, int(ws))
        assert len(a) == nw and int(a[-1][0, 0]) == start_last and int(a[-1][-1, 0]) == end_last
        if nw > 1:
            assert int(a[1][0, 0]) == start1

"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (from /root/reference)
under the TF->torch shim in oracle/refshim.  Run in the build container only:

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The committed .npz files are what travels to the GPU box; nothing at
test/bench time reads /root/reference.  Gradients are torch-autograd gradients of the reference's
graph-building code w.r.t. the GPflow *free* state (what tf.gradients feeds the optimiser).
"""
import json
import os
import sys
import numpy as np
import torch

from .refshim import loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrs)
    print('wrote', name, {k: np.asarray(v).shape for k, v in arrs.items()})


def _np(t):
    return t.detach().numpy().copy()


def harmonic(Q, f0, rng):
    e = 1. / (np.arange(1, Q + 1) ** 2)
    e = e / e.sum()
    f = f0 * np.arange(1, Q + 1) * (1. + 1e-3 * rng.standard_normal(Q))
    return e, f


def golden_kernels(ns, rng):
    for tag, t0 in (('t0', 0.0), ('t10', 10.0), ('t240', 240.0)):
        N, M, Q = 96, 24, 5
        x = (t0 + np.arange(N) / 16000.).reshape(-1, 1)
        z = x[::4].copy()
        e, f = harmonic(Q, 261.6255653005986, rng)
        var, ls = 1.3, 0.05
        k = ns.mm.MercerMatern12sm(1, energy=e, frequency=f, variance=var, lengthscales=ls)
        Kzx = _np(k.K(torch.as_tensor(z), torch.as_tensor(x)))
        Kzz = _np(k.K(torch.as_tensor(z)))
        Kd = _np(k.Kdiag(torch.as_tensor(x)))
        phi = _np(k.phi_features(torch.as_tensor(x)))
        k2 = ns.mm.Matern12sm(1, variance=var, lengthscales=ls, energy=e, frequency=f)
        K2zx = _np(k2.K(torch.as_tensor(z), torch.as_tensor(x)))
        K2zz = _np(k2.K(torch.as_tensor(z)))
        K2d = _np(k2.Kdiag(torch.as_tensor(x)))
        _save('kernels_' + tag, x=x, z=z, energy=e, frequency=f, variance=var, lengthscales=ls,
              mercer_Kzx=Kzx, mercer_Kzz=Kzz, mercer_Kdiag=Kd, mercer_phi=phi,
              diff_Kzx=K2zx, diff_Kzz=K2zz, diff_Kdiag=K2d)


def golden_nonlin(ns, rng):
    x = rng.standard_normal(64) * 3 + 2
    _save('nonlin', x=x, logistic=ns.methods.logistic(x), softplus=ns.methods.softplus(x),
          gaussfun=ns.methods.gaussfun(x),
          logistic_tf=_np(ns.methods.logistic_tf(torch.as_tensor(x))),
          softplus_tf=_np(ns.methods.softplus_tf(torch.as_tensor(x))),
          gaussfun_tf=_np(ns.methods.gaussfun_tf(torch.as_tensor(x))),
          midi2freq60=ns.methods.midi2freq(60), f0_M60=np.asarray(ns.methods.find_ideal_f0(['011PFNOF_M60_train.wav'])))


def golden_likelihoods(ns, rng):
    for P in (1, 3):
        n = 40
        Fmu = rng.standard_normal((n, 2 * P)) * 2 + 1.5
        Fvar = np.exp(rng.standard_normal((n, 2 * P)) * 1.5 - 1.0)
        Y = rng.standard_normal((n, 1))
        out = {}
        for name, fn in (('logistic', ns.methods.logistic_tf), ('softplus', ns.methods.softplus_tf),
                         ('gauss', ns.methods.gaussfun_tf)):
            lik = ns.likelihoods.MpdLik(nlinfun=fn, num_sources=P)
            lik.variance = 0.37
            tFmu = torch.tensor(Fmu, requires_grad=True)
            tFvar = torch.tensor(Fvar, requires_grad=True)
            free = lik.raw('variance').free
            ve = lik.variational_expectations(tFmu, tFvar, torch.as_tensor(Y))
            g = torch.autograd.grad(ve.sum(), [tFmu, tFvar, free])
            out['ve_' + name] = _np(ve)
            out['dFmu_' + name] = g[0].numpy()
            out['dFvar_' + name] = g[1].numpy()
            out['dfree_noise_' + name] = g[2].numpy()
            F = rng.standard_normal((n, 2 * P))
            out['F_' + name] = F
            out['logp_' + name] = _np(lik.logp(torch.as_tensor(F), torch.as_tensor(Y)))
        _save('mpdlik_P%d' % P, Fmu=Fmu, Fvar=Fvar, Y=Y, noise_var=0.37, **out)
    # single-source ModLik (columns are [f, g])
    n = 40
    Fmu = rng.standard_normal((n, 2)) * 2 + 1.5
    Fvar = np.exp(rng.standard_normal((n, 2)) - 1.0)
    Y = rng.standard_normal((n, 1))
    lik = ns.likelihoods.ModLik(ns.methods.logistic_tf)
    lik.variance = 0.8
    ve = lik.variational_expectations(torch.as_tensor(Fmu), torch.as_tensor(Fvar), torch.as_tensor(Y))
    _save('modlik', Fmu=Fmu, Fvar=Fvar, Y=Y, noise_var=0.8, ve=_np(ve))


def _signal(x, f0s, rng):
    y = np.zeros_like(x)
    for f0 in f0s:
        for q in range(1, 4):
            y += np.sin(2 * np.pi * q * f0 * x + rng.uniform(0, 2 * np.pi)) / q ** 2
    y += 0.01 * rng.standard_normal(x.shape)
    return y / np.max(np.abs(y))


def golden_sgprss(ns, rng):
    N, M, Q, P = 160, 20, 4, 3
    for tag, t0 in (('t0', 0.0), ('t10', 10.0)):
        x = (t0 + np.arange(N) / 16000.).reshape(-1, 1)
        z = x[::N // M].copy()
        f0s = [ns.methods.midi2freq(m) for m in (60, 64, 67)]
        y = _signal(x, f0s, rng)
        es, fs = zip(*[harmonic(Q, f0, rng) for f0 in f0s])
        variances = [0.9, 1.4, 0.6]
        ls = [0.1, 0.07, 0.2]
        for reg in (False, True):
            kerns = ns.init_kernels.init_kern_com(P, [np.asarray(l) for l in ls], list(es), list(fs), len_fixed=False)
            for k, v in zip(kerns, variances):
                k.variance = v
            kern = np.sum(kerns)
            m = ns.sgpr_ss.SGPRSS(X=x, Y=y, kern=kern, Z=z, reg=reg)
            m.likelihood.variance = 0.05
            fval, grads = m.objective_and_grads()
            xnew = x[::3].copy()
            mf, vf = m.predict_f(xnew)
            ms, vs = m.predict_s(xnew)
            names = sorted(grads)
            _save('sgprss_%s_reg%d' % (tag, int(reg)), x=x, y=y, z=z, xnew=xnew,
                  energy=np.asarray(es), frequency=np.asarray(fs), variance=np.asarray(variances),
                  lengthscales=np.asarray(ls), noise_var=0.05, neg_bound=fval,
                  grad_names=json.dumps(names), grads=np.concatenate([grads[n].ravel() for n in names]),
                  predict_f_mean=mf, predict_f_var=vf, predict_s_mean=np.asarray(ms), predict_s_var=np.asarray(vs))


def golden_pdgp(ns, rng):
    N, M, Q = 120, 15, 4
    for P, whiten, zfree in [(p_, w_, False) for p_ in (1, 2) for w_ in (True, False)] + [(2, True, True), (2, False, True)]:
        if True:
            r = np.random.default_rng(4242 + int(whiten)) if zfree else rng     # keeps the main stream unchanged
            x = (2.0 + np.arange(N) / 16000.).reshape(-1, 1)
            z = x[::N // M].copy()
            f0s = [ns.methods.midi2freq(m) for m in (60, 67)][:P]
            y = _signal(x, f0s, r)
            es, fs = zip(*[harmonic(Q, f0, r) for f0 in f0s])
            ls = [0.05, 0.08][:P]
            kern_com = ns.init_kernels.init_kern_com(P, [np.asarray(l) for l in ls], list(es), list(fs), len_fixed=False)
            kern_act = ns.init_kernels.init_kern_act(P)
            for i, k in enumerate(kern_act):
                k.lengthscales = 0.002 * (i + 1)      # resolvable at this tiny window length
            zz = [[z.copy() for _ in range(P)], [z.copy() for _ in range(P)]]
            m = ns.pdgp.Pdgp(x, y, zz, [kern_act, kern_com], whiten=whiten)
            q_mu_a = [0.3 * r.standard_normal((M, 1)) + 1.0 for _ in range(P)]
            q_mu_c = [0.3 * r.standard_normal((M, 1)) for _ in range(P)]
            q_sq_a = [(np.eye(M) * 0.5 + 0.05 * r.standard_normal((M, M)))[:, :, None] for _ in range(P)]
            q_sq_c = [(np.eye(M) * 0.7 + 0.05 * r.standard_normal((M, M)))[:, :, None] for _ in range(P)]
            for i in range(P):
                m.q_mu_act.raw_item(i).set(q_mu_a[i]); m.q_mu_com.raw_item(i).set(q_mu_c[i])
                m.q_sqrt_act.raw_item(i).set(q_sq_a[i]); m.q_sqrt_com.raw_item(i).set(q_sq_c[i])
                if not zfree:
                    m.za.raw_item(i).fixed = True; m.zc.raw_item(i).fixed = True      # demo-modgp.py:40-41
            m.likelihood.variance = 0.02
            fval, grads = m.objective_and_grads()
            kl = float(m.build_prior_kl())
            xnew = x[::2].copy()
            ma, va, mc, vc, msrc = m.predict_act_n_com(xnew)
            names = sorted(grads)
            _save('pdgp_P%d_whiten%d%s' % (P, int(whiten), '_zfree' if zfree else ''), x=x, y=y, z=z, xnew=xnew,
                  energy=np.asarray(es), frequency=np.asarray(fs), lengthscales_com=np.asarray(ls),
                  variance_com=np.ones(P), variance_act=3.5 * np.ones(P),
                  lengthscales_act=np.asarray([0.002 * (i + 1) for i in range(P)]),
                  q_mu_act=np.asarray(q_mu_a), q_mu_com=np.asarray(q_mu_c),
                  q_sqrt_act=np.asarray(q_sq_a), q_sqrt_com=np.asarray(q_sq_c), noise_var=0.02,
                  neg_elbo=fval, prior_kl=kl, grad_names=json.dumps(names),
                  grad_sizes=np.asarray([grads[n].size for n in names]),
                  grads=np.concatenate([grads[n].ravel() for n in names]),
                  mean_act=np.asarray(ma), var_act=np.asarray(va), mean_com=np.asarray(mc), var_com=np.asarray(vc),
                  mean_source=np.asarray(msrc))


def golden_windows(ns, rng):
    n, ws = 1507, 201
    x = np.linspace(0., (n - 1.) / 16000., n).reshape(-1, 1)
    y = rng.standard_normal((n, 1))
    xw, yw = ns.window_overlap.windowed(x, y, ws)
    nmerged = (ws - 1) // 2 * (len(xw) - 1) + ws
    mm = ns.window_overlap.merged_mean([w.copy() for w in yw], ws, nmerged)
    mv = ns.window_overlap.merged_variance([np.abs(w) for w in yw], ws, nmerged)
    mx = ns.window_overlap.merged_x([w.copy() for w in xw], ws)
    xs, ys = ns.window_overlap.segmented(x, y, window_size=300, aug=False)
    xa, ya = ns.window_overlap.segmented(x, y, window_size=300, aug=True)
    _save('windows', x=x, y=y, ws=ws, xw=np.asarray(xw), yw=np.asarray(yw), merged_mean=mm, merged_variance=mv,
          merged_x=mx, seg_x=np.asarray(xs), seg_y=np.asarray(ys), aug_x=np.asarray(xa), aug_y=np.asarray(ya))
    # geometry table: the Python-2 integer arithmetic that defines window counts / starts
    rows = []
    for n_, ws_ in ((224001, 2001), (3840000, 2001), (32000, 1601), (5000, 401), (2001, 2001), (4000, 801)):
        xx = np.arange(n_, dtype=np.float64)
        a, _ = ns.window_overlap.windowed(xx, xx, ws_)
        rows.append([n_, ws_, len(a), int(a[1][0, 0]) if len(a) > 1 else -1, int(a[-1][0, 0]), int(a[-1][-1, 0])])
    _save('window_geometry', table=np.asarray(rows, dtype=np.int64))


def golden_init_models(ns, rng):
    """init_liv / init_iv through the reference's own gpitch/init_models.py (needs gpitch/kernels.py importable under
    the shim).  The shipped wav reproduces the notebook's known answer (109 inducing points); only the OUTPUTS of that
    case are stored (the wav is not copied) together with a synthetic case whose inputs are stored too."""
    import hashlib
    import sys as _sys
    from scipy.io import wavfile
    k = loader._load('gpitch.kernels', 'gpitch/kernels.py')
    _sys.modules['gpitch'].kernels = k
    im = loader._load('gpitch.init_models', 'gpitch/init_models.py')
    wav = os.path.join(loader.REF, 'demos', 'data', '011PFNOF_M60_train.wav')
    fs, y = wavfile.read(wav)
    y = y.astype(np.float64).reshape(-1, 1)
    x = np.linspace(0., (y.size - 1.) / fs, y.size).reshape(-1, 1)
    z, yf = im.init_liv(x, y, win_size=31, thres=0.033, dec=9)           # demo_modgp-real-audio.ipynb cell 5 -> 109
    rng2 = np.random.default_rng(3)
    n = 4000
    xs = np.arange(n).reshape(-1, 1) / 16000.
    ys = np.sin(2 * np.pi * 220 * xs) * np.exp(-((xs - 0.12) / 0.05) ** 2) + 0.02 * rng2.standard_normal((n, 1))
    zs, yfs = im.init_liv(xs, ys, num_sources=2, win_size=9, thres=0.05, dec=2)
    ziv = im.init_iv(xs, 2, 400, 800, 16000)
    _save('init_models', xs=xs, ys=ys, liv_z=zs[0][0], liv_y=yfs, iv_za=ziv[0][0], iv_zc=ziv[1][0],
          wav_liv_z=z[0][0], wav_liv_y=yf, wav_sha1=hashlib.sha1(open(wav, 'rb').read()).hexdigest())


def golden_legacy_kernels(ns):
    """Matern32sm / Matern32sml of gpitch/kernels.py:204-318 (legacy init_models only)."""
    import sys as _sys
    k = _sys.modules.get('gpitch.kernels') or loader._load('gpitch.kernels', 'gpitch/kernels.py')
    rng = np.random.default_rng(31)
    N, M, Q = 90, 18, 4
    x = (0.5 + np.arange(N) / 16000.).reshape(-1, 1)
    z = x[::5].copy()
    freqs = 220.0 * (1. + np.arange(Q)) * (1 + 1e-3 * rng.standard_normal(Q))
    var_sm = rng.uniform(0.02, 0.2, (Q, 1)); var_sml = rng.uniform(0.05, 0.9, (Q, 1))
    ls_sml = rng.uniform(0.005, 0.05, (Q, 1))
    ksm = k.Matern32sm(1, Q, lengthscales=0.02, variances=var_sm, frequencies=freqs)
    ksml = k.Matern32sml(1, Q, lengthscales=ls_sml, variances=var_sml, frequencies=freqs)
    tx, tz = torch.as_tensor(x), torch.as_tensor(z)
    val = lambda pl: np.asarray([np.squeeze(p.value) for p in pl])
    _save('legacy_kernels', x=x, z=z,
          sm_lengthscales=np.squeeze(ksm.raw('lengthscales').value), sm_variance=val(ksm.variance._list if hasattr(ksm.variance, '_list') else ksm.variance),
          sm_frequency=val(ksm.frequency._list if hasattr(ksm.frequency, '_list') else ksm.frequency),
          sm_Kzx=_np(ksm.K(tz, tx)), sm_Kzz=_np(ksm.K(tz)), sm_Kdiag=_np(ksm.Kdiag(tx)),
          sml_lengthscales=val(ksml.lengthscales._list if hasattr(ksml.lengthscales, '_list') else ksml.lengthscales),
          sml_variance=val(ksml.variance._list if hasattr(ksml.variance, '_list') else ksml.variance),
          sml_frequency=val(ksml.frequency._list if hasattr(ksml.frequency, '_list') else ksml.frequency),
          sml_Kzx=_np(ksml.K(tz, tx)), sml_Kzz=_np(ksml.K(tz)), sml_Kdiag=_np(ksml.Kdiag(tx)))


def golden_kernelfit(ns):
    """gpitch/kernelfit.py profile functions (pure NumPy in the reference; executed from its own source)."""
    import sys as _sys
    kf = _sys.modules.get('gpitch.kernelfit') or loader._load('gpitch.kernelfit', 'gpitch/kernelfit.py')
    rng = np.random.default_rng(77)
    n, m, fs = 400, 4, 16000.
    x = np.linspace(0., (n - 1.) / fs, n).reshape(-1, 1)
    p = np.hstack(([0.3, 0.012], rng.uniform(0.05, 0.3, m), 220. * (1 + np.arange(m)) * (1 + 1e-3 * rng.standard_normal(m))))
    p[3] *= -1.0                                   # the reference takes sqrt(p * p): signs must not matter
    y = kf.approximate_kernel(p * (1 + 0.05 * rng.standard_normal(p.size)), x) + 0.01 * rng.standard_normal((n, 1))
    pg = np.array([0.4, 0.01, 440., 0.2, 0.02, 880.])
    _save('kernelfit', x=x, p=p, y=y, approx=kf.approximate_kernel(p, x), loss=kf.loss_func(p, x, y),
          gabor_p=pg, gabor_sum=kf.func(x.reshape(-1), *pg))


def golden_large(ns):
    """Round-2 additions (own random streams, so the files above stay byte-identical): SGPRSS at the full C1 shape
    (N = 1600, M = 200, P = 3, Q = 10, BASELINE configs[0]) and SGPRSS / Pdgp cases at a late absolute time stamp
    (t = 240 s: the end of a 4-minute track, where GPflow's distance-by-expansion is noisiest)."""
    for tag, N, M, Q, t0 in (('c1', 1600, 200, 10, 0.0), ('t240', 160, 20, 4, 240.0)):
        rng = np.random.default_rng(777 + N)
        P = 3
        x = (t0 + np.arange(N) / 16000.).reshape(-1, 1)
        z = x[::N // M].copy()
        f0s = [ns.methods.midi2freq(m) for m in (60, 64, 67)]
        y = _signal(x, f0s, rng)
        es, fs = zip(*[harmonic(Q, f0, rng) for f0 in f0s])
        variances = [0.9, 1.4, 0.6]
        ls = [0.1, 0.07, 0.2]
        kerns = ns.init_kernels.init_kern_com(P, [np.asarray(l) for l in ls], list(es), list(fs), len_fixed=False)
        for k, v in zip(kerns, variances):
            k.variance = v
        m = ns.sgpr_ss.SGPRSS(X=x, Y=y, kern=np.sum(kerns), Z=z, reg=False)
        m.likelihood.variance = 0.05
        fval, grads = m.objective_and_grads()
        xnew = x[::8].copy() if N > 1000 else x[::3].copy()
        mf, vf = m.predict_f(xnew)
        ms, vs = m.predict_s(xnew)
        names = sorted(grads)
        _save('sgprss_%s_reg0' % tag, x=x, y=y, z=z, xnew=xnew, energy=np.asarray(es), frequency=np.asarray(fs),
              variance=np.asarray(variances), lengthscales=np.asarray(ls), noise_var=0.05, neg_bound=fval,
              grad_names=json.dumps(names), grads=np.concatenate([grads[n].ravel() for n in names]),
              predict_f_mean=mf, predict_f_var=vf, predict_s_mean=np.asarray(ms), predict_s_var=np.asarray(vs))
    # Pdgp, whitened, fixed inducing inputs, t = 240 s
    N, M, Q, P = 120, 15, 4, 2
    r = np.random.default_rng(909)
    x = (240.0 + np.arange(N) / 16000.).reshape(-1, 1)
    z = x[::N // M].copy()
    f0s = [ns.methods.midi2freq(m) for m in (60, 67)]
    y = _signal(x, f0s, r)
    es, fs = zip(*[harmonic(Q, f0, r) for f0 in f0s])
    ls = [0.05, 0.08]
    kern_com = ns.init_kernels.init_kern_com(P, [np.asarray(l) for l in ls], list(es), list(fs), len_fixed=False)
    kern_act = ns.init_kernels.init_kern_act(P)
    for i, k in enumerate(kern_act):
        k.lengthscales = 0.002 * (i + 1)
    zz = [[z.copy() for _ in range(P)], [z.copy() for _ in range(P)]]
    m = ns.pdgp.Pdgp(x, y, zz, [kern_act, kern_com], whiten=True)
    q_mu_a = [0.3 * r.standard_normal((M, 1)) + 1.0 for _ in range(P)]
    q_mu_c = [0.3 * r.standard_normal((M, 1)) for _ in range(P)]
    q_sq_a = [(np.eye(M) * 0.5 + 0.05 * r.standard_normal((M, M)))[:, :, None] for _ in range(P)]
    q_sq_c = [(np.eye(M) * 0.7 + 0.05 * r.standard_normal((M, M)))[:, :, None] for _ in range(P)]
    for i in range(P):
        m.q_mu_act.raw_item(i).set(q_mu_a[i]); m.q_mu_com.raw_item(i).set(q_mu_c[i])
        m.q_sqrt_act.raw_item(i).set(q_sq_a[i]); m.q_sqrt_com.raw_item(i).set(q_sq_c[i])
        m.za.raw_item(i).fixed = True; m.zc.raw_item(i).fixed = True
    m.likelihood.variance = 0.02
    fval, grads = m.objective_and_grads()
    kl = float(m.build_prior_kl())
    xnew = x[::2].copy()
    ma, va, mc, vc, msrc = m.predict_act_n_com(xnew)
    names = sorted(grads)
    _save('pdgp_P2_whiten1_t240', x=x, y=y, z=z, xnew=xnew, energy=np.asarray(es), frequency=np.asarray(fs),
          lengthscales_com=np.asarray(ls), variance_com=np.ones(P), variance_act=3.5 * np.ones(P),
          lengthscales_act=np.asarray([0.002 * (i + 1) for i in range(P)]),
          q_mu_act=np.asarray(q_mu_a), q_mu_com=np.asarray(q_mu_c), q_sqrt_act=np.asarray(q_sq_a),
          q_sqrt_com=np.asarray(q_sq_c), noise_var=0.02, neg_elbo=fval, prior_kl=kl, grad_names=json.dumps(names),
          grad_sizes=np.asarray([grads[n].size for n in names]), grads=np.concatenate([grads[n].ravel() for n in names]),
          mean_act=np.asarray(ma), var_act=np.asarray(va), mean_com=np.asarray(mc), var_com=np.asarray(vc),
          mean_source=np.asarray(msrc))


def main():
    ns = loader.load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'large':       # round-2 files only
        return golden_large(ns)
    rng = np.random.default_rng(20261018)
    golden_kernels(ns, rng)
    golden_nonlin(ns, rng)
    golden_likelihoods(ns, rng)
    golden_sgprss(ns, rng)
    golden_pdgp(ns, rng)
    golden_windows(ns, rng)
    golden_init_models(ns, rng)
    golden_legacy_kernels(ns)
    golden_kernelfit(ns)
    golden_large(ns)


if __name__ == '__main__':
    sys.exit(main())

"""``Pdgp`` -- pitch detection with a modulated GP (SVGP, P activation GPs x P component GPs), with the constructor,
parameter names and predict methods of gpitch/pdgp.py:48-208.  One instance = one window; the arithmetic runs in
the window-batched CUDA engine (gpitch_b200/batched.py, W = 1)."""
import numpy as np
import torch

from . import train
from .batched import BatchedPdgp, GraphedEvaluation
from .likelihoods import MpdLik
from .methods import logistic_tf, nlin_name
from .param import Param, ParamList, Parameterized


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class _Minibatch(object):
    """gpflow.minibatch.MinibatchData stand-in: index stream from RandomState(0) (pdgp.py:76-77).  x and y use
    identically seeded generators, so one stream serves both.  [GPflow-0.5, recalled]: indices are drawn with
    replacement when minibatch/total is small, else a prefix of a permutation; the 0.5 threshold is uncertain."""
    def __init__(self, total, size, seed=0):
        self.total, self.size, self.rng = total, size, np.random.RandomState(seed)

    def next(self):
        if self.size >= self.total:
            return np.arange(self.total)
        if float(self.size) / self.total > 0.5:
            return self.rng.permutation(self.total)[:self.size]
        return self.rng.randint(self.total, size=self.size)


class Pdgp(Parameterized):
    def __init__(self, x, y, z, kern, whiten=True, minibatch_size=None, nlinfun=logistic_tf):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if minibatch_size is None:
            minibatch_size = x.shape[0]
        self.minibatch_size = minibatch_size
        self.num_data = x.shape[0]
        self.num_sources = len(kern[0])
        self.whiten = whiten
        self.nlinfun = nlinfun
        self.likelihood = MpdLik(nlinfun=nlinfun, num_sources=self.num_sources)
        self._x, self._y = x.reshape(-1), y.reshape(-1)
        self._mb = _Minibatch(self.num_data, minibatch_size)
        self.kern_act = ParamList(kern[0])
        self.kern_com = ParamList(kern[1])
        P = self.num_sources
        self.num_inducing_a = [np.asarray(z[0][i]).size for i in range(P)]
        self.num_inducing_c = [np.asarray(z[1][i]).size for i in range(P)]
        if len(set(self.num_inducing_a)) != 1 or len(set(self.num_inducing_c)) != 1:
            raise NotImplementedError('all activation (resp. component) GPs of a window share one inducing count')
        self.za = ParamList([Param(np.asarray(z[0][i]).copy()) for i in range(P)])
        self.zc = ParamList([Param(np.asarray(z[1][i]).copy()) for i in range(P)])
        self.q_mu_act = ParamList([Param(np.zeros(np.asarray(z[0][i]).shape)) for i in range(P)])
        self.q_mu_com = ParamList([Param(np.zeros(np.asarray(z[1][i]).shape)) for i in range(P)])
        self.q_sqrt_act = ParamList([Param(np.eye(self.num_inducing_a[i])[:, :, None]) for i in range(P)])
        self.q_sqrt_com = ParamList([Param(np.eye(self.num_inducing_c[i])[:, :, None]) for i in range(P)])

    # ------------------------------------------------------------------ packing
    def _pack(self):
        P = self.num_sources
        Q = max(k.num_q() for k in self.kern_com)
        d = {'act_hyp': np.stack([k.hyper_row(0) for k in self.kern_act])[None],
             'com_hyp': np.stack([k.hyper_row(Q) for k in self.kern_com])[None],
             'q_mu_act': np.stack([p.value.reshape(-1) for p in self.q_mu_act])[None],
             'q_mu_com': np.stack([p.value.reshape(-1) for p in self.q_mu_com])[None],
             'q_sqrt_act': np.stack([p.value[:, :, 0] for p in self.q_sqrt_act])[None],
             'q_sqrt_com': np.stack([p.value[:, :, 0] for p in self.q_sqrt_com])[None],
             'noise': np.array([float(np.squeeze(self.likelihood.variance.value))])}
        return {k: _dev(v) for k, v in d.items()}, Q

    use_cuda_graph = True      # replay one captured graph per evaluation (single windows are launch-bound)

    def _engine(self, idx=None):
        """Persistent W = 1 engine (per data shape); minibatches / new data update its buffers in place."""
        x = self._x if idx is None else self._x[idx]
        y = self._y if idx is None else self._y[idx]
        za = np.stack([p.value.reshape(-1) for p in self.za])[None]
        zc = np.stack([p.value.reshape(-1) for p in self.zc])[None]
        kc = self.kern_com[0]
        # za / zc are Params like in the reference (pdgp.py:80-85); demo-modgp.py:40-41 fixes them.  The inducing-point
        # gradient kernels only run while at least one of them is free.
        train_z = not all(p.fixed for p in list(self.za) + list(self.zc))
        key = (x.shape, za.shape, zc.shape, nlin_name(self.nlinfun), kc.distance_mode, kc.kind, self.whiten, kc.num_q(),
               train_z)
        cache = self.__dict__.get('_eng_cache')
        if cache is None or cache[0] != key:
            eng = BatchedPdgp(_dev(x[None]), _dev(y[None]), _dev(za), _dev(zc), nlin=nlin_name(self.nlinfun),
                              mode=kc.distance_mode, kind_com=kc.kind, whiten=self.whiten, train_z=train_z)
            object.__setattr__(self, '_eng_cache', (key, eng, {}))
        else:
            cache[1].set_data(_dev(x[None]), _dev(y[None]), _dev(za), _dev(zc))
        return self.__dict__['_eng_cache'][1]

    def _graphed_elbo(self, eng, d):
        graphs = self.__dict__['_eng_cache'][2]
        # The conditional() formulation chosen by the engine's 'auto' certificate is baked into the captured graph: it is
        # re-certified eagerly every GFORM_RECHECK replays (hyper-parameters move during optimisation) and the graph is
        # re-captured when the choice changed.
        graphs['n'] = graphs.get('n', 0) + 1
        if 'elbo' in graphs and graphs['n'] % eng.GFORM_RECHECK == 0 and eng.recertify_gform(d['act_hyp'], d['com_hyp']):
            del graphs['elbo']
        if 'elbo' not in graphs:
            nd = self.num_data
            graphs['elbo'] = GraphedEvaluation(
                lambda **p: eng.elbo(*[p[k] for k in BatchedPdgp.NAMES], need_grad=True, num_data=nd) + (eng.last_info,), d)
        return graphs['elbo'](**d)

    # ------------------------------------------------------------------ objective
    def build_prior_kl(self):
        """pdgp.py:113-131: sum of the 2P gauss_kl terms (whitened, or against K(z) + jitter I when whiten=False)."""
        from . import _lib
        from .functions import KernelMatrix, Unwhiten
        d, _ = self._pack()
        eng = self._engine()
        kl = 0.0
        with torch.no_grad():
            for grp, kind, z in (('act', 'matern32', eng.za), ('com', eng.kind_com, eng.zc)):
                mu, sq = d['q_mu_' + grp][0].contiguous(), d['q_sqrt_' + grp][0].contiguous()
                if not self.whiten:
                    hyp = d[grp + '_hyp'][0].unsqueeze(1).contiguous()
                    Kmm = KernelMatrix.apply(hyp, z[0].contiguous(), z[0].contiguous(), kind, eng.mode, eng.jitter, False)
                    mu, sq = Unwhiten.apply(mu, sq, Kmm)[:2]
                kl += float(_lib.gauss_kl_white(mu.contiguous(), sq.contiguous(), need_grad=False)[0].sum())
        return kl

    def build_likelihood(self):
        """pdgp.py:133-170: ELBO at the current parameters (draws the next minibatch if one is configured)."""
        d, _ = self._pack()
        eng = self._engine(self._mb.next() if self.minibatch_size < self.num_data else None)
        e, _ = eng.elbo(*[d[k] for k in BatchedPdgp.NAMES], need_grad=False, num_data=self.num_data)
        return float(e[0])

    compute_log_likelihood = build_likelihood

    def _objective(self, x):
        self.set_state(x)
        d, Q = self._pack()
        eng = self._engine(self._mb.next() if self.minibatch_size < self.num_data else None)
        if self.use_cuda_graph:
            e, g, info = self._graphed_elbo(eng, d)     # status captured as a static graph output (eager calls rebind last_info)
        else:
            e, g = eng.elbo(*[d[k] for k in BatchedPdgp.NAMES], need_grad=True, num_data=self.num_data)
            info = eng.last_info
        if int(info.abs().max()) != 0:
            return np.inf, np.zeros_like(np.asarray(x, dtype=np.float64))
        g = {k: v[0].cpu().numpy() for k, v in g.items()}
        grads = {id(self.likelihood.variance): g['noise']}
        for i in range(self.num_sources):
            ka, kc = self.kern_act[i], self.kern_com[i]
            grads[id(ka.variance)] = g['act_hyp'][i, 0]
            grads[id(ka.lengthscales)] = g['act_hyp'][i, 1]
            grads[id(kc.variance)] = g['com_hyp'][i, 0]
            grads[id(kc.lengthscales)] = g['com_hyp'][i, 1]
            for q in range(kc.num_q()):
                grads[id(kc.energy[q])] = g['com_hyp'][i, 2 + q]
                grads[id(kc.frequency[q])] = g['com_hyp'][i, 2 + Q + q]
            grads[id(self.q_mu_act[i])] = g['q_mu_act'][i]
            grads[id(self.q_mu_com[i])] = g['q_mu_com'][i]
            grads[id(self.q_sqrt_act[i])] = g['q_sqrt_act'][i]
            grads[id(self.q_sqrt_com[i])] = g['q_sqrt_com'][i]
            if 'za' in g:
                grads[id(self.za[i])] = g['za'][i]
                grads[id(self.zc[i])] = g['zc'][i]
        out = [np.ravel(grads[id(p)]) * p.chain() for _, p in self.free_params()]
        return -float(e[0]), -np.concatenate(out)

    def optimize(self, method='L-BFGS-B', tol=None, callback=None, maxiter=1000, **kw):
        return train.optimize(self, method=method, tol=tol, callback=callback, maxiter=maxiter, **kw)

    # ------------------------------------------------------------------ predictions (pdgp.py:172-208)
    def _predict(self, xnew):
        d, _ = self._pack()
        eng = self._engine()
        out = eng.predict(_dev(np.asarray(xnew, dtype=np.float64).reshape(1, -1)), *[d[k] for k in BatchedPdgp.NAMES[:6]])
        P = self.num_sources
        return [[t[0, i].cpu().numpy().reshape(-1, 1) for i in range(P)] for t in out]

    def predict_act(self, xnew):
        ma, va, _, _, _ = self._predict(xnew)
        return ma, va

    def predict_com(self, xnew):
        _, _, mc, vc, _ = self._predict(xnew)
        return mc, vc

    def predict_act_n_com(self, xnew):
        return tuple(self._predict(xnew))

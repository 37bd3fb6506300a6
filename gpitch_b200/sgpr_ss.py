"""``SGPRSS`` -- sparse GP regression for source separation, with the constructor / methods / assignable
attributes of gpitch/sgpr_ss.py:10-114 (subclass of GPflow SGPR in the reference).  One instance = one audio
window; it drives the window-batched CUDA engine (gpitch_b200/batched.py) with W = 1.  For many windows at once use
``gpitch_b200.batched.BatchedSGPR`` directly (what `SoSp.optimize` / `AMT.optimize` loop over in the reference)."""
import numpy as np
import torch

from . import train
from .batched import BatchedSGPR, GraphedEvaluation
from .kernels import Add
from .likelihoods import Gaussian
from .param import DataHolder, Parameterized


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class SGPRSS(Parameterized):
    def __init__(self, X, Y, kern, Z, mean_function=None, reg=False):
        if mean_function is not None:
            raise NotImplementedError('gpitch only ever uses the Zero mean function')
        self.X = DataHolder(X, on_shape_change='pass')
        self.Y = DataHolder(Y, on_shape_change='pass')
        self.Z = DataHolder(Z, on_shape_change='pass')          # sgpr_ss.py:26: Z is data, not a Param
        self.kern = kern
        self.likelihood = Gaussian()
        self.reg = reg
        self.num_latent = np.asarray(Y).shape[1]
        if self.num_latent != 1:
            raise NotImplementedError('gpitch audio windows have one output column')

    # ------------------------------------------------------------------ packing
    def _components(self):
        return self.kern.components()

    def _hyp(self):
        comps = self._components()
        Q = max(c.num_q() for c in comps)
        return np.stack([c.hyper_row(Q) for c in comps])[None], Q

    use_cuda_graph = True      # replay one captured graph per evaluation (single windows are launch-bound)

    def _engine(self):
        """Persistent W = 1 engine; window swaps (model.X = ..., separation.py:266-268) update its buffers in place."""
        comps = self._components()
        kind = comps[0].kind
        if any(c.kind != kind for c in comps):
            raise NotImplementedError('Add of mixed kernel kinds is not on the gpitch hot path')
        x, y, z = (_dev(a.value.reshape(1, -1)) for a in (self.X, self.Y, self.Z))
        key = (x.shape, z.shape, kind, comps[0].distance_mode, self.reg, len(comps), max(c.num_q() for c in comps))
        cache = self.__dict__.get('_eng_cache')
        if cache is None or cache[0] != key:
            eng = BatchedSGPR(x, y, z, kind=kind, mode=comps[0].distance_mode, reg=self.reg)
            object.__setattr__(self, '_eng_cache', (key, eng, {}))
        else:
            cache[1].set_data(x, y, z)
        return self.__dict__['_eng_cache'][1]

    def _graphed_bound(self, eng, hyp, noise):
        graphs = self.__dict__['_eng_cache'][2]
        if 'bound' not in graphs:
            # the Cholesky status is part of the graph's static outputs: eager engine calls (predict_f / predict_s between
            # two windows of the SoSp loop) rebind eng.last_info, a replay writes into the tensor captured here
            graphs['bound'] = GraphedEvaluation(lambda hyp, noise: eng.bound(hyp, noise, need_grad=True) + (eng.last_info,),
                                                {'hyp': hyp, 'noise': noise})
        return graphs['bound'](hyp=hyp, noise=noise)

    def _noise(self):
        return _dev([float(np.squeeze(self.likelihood.variance.value))])

    # ------------------------------------------------------------------ objective
    def build_likelihood(self):
        """Value of the collapsed bound (sgpr_ss.py:29-71) at the current parameters."""
        hyp, _ = self._hyp()
        eng = self._engine()
        b, _ = eng.bound(_dev(hyp), self._noise(), need_grad=False)
        self._check(eng)
        return float(b[0])

    compute_log_likelihood = build_likelihood

    def _check(self, eng):
        if int(eng.last_info.abs().max()) != 0:
            raise FloatingPointError('Cholesky failed (matrix not positive definite): info=%s'
                                     % eng.last_info.cpu().tolist())

    def _objective(self, x):
        """GPflow Model._objective: free state -> (-bound, -d bound / d free state)."""
        self.set_state(x)
        hyp, Q = self._hyp()
        eng = self._engine()
        if self.use_cuda_graph:
            b, g, info = self._graphed_bound(eng, _dev(hyp), self._noise())
        else:
            b, g = eng.bound(_dev(hyp), self._noise(), need_grad=True)
            info = eng.last_info
        if int(info.abs().max()) != 0:                   # failed window: -inf bound, zero gradient
            return np.inf, np.zeros_like(np.asarray(x, dtype=np.float64))
        gh = g['hyp'][0].cpu().numpy()
        gn = float(g['noise'][0])
        grads = {}
        for i, c in enumerate(self._components()):
            grads[id(c.variance)] = grads.get(id(c.variance), 0.0) + gh[i, 0]
            grads[id(c.lengthscales)] = grads.get(id(c.lengthscales), 0.0) + gh[i, 1]
            for q in range(c.num_q()):
                grads[id(c.energy[q])] = gh[i, 2 + q]
                grads[id(c.frequency[q])] = gh[i, 2 + Q + q]
        grads[id(self.likelihood.variance)] = gn
        out = []
        for _, p in self.free_params():
            out.append(np.atleast_1d(grads.get(id(p), 0.0)) * p.chain())
        return -float(b[0]), -np.concatenate(out) if out else np.zeros(0)

    def optimize(self, method='L-BFGS-B', tol=None, callback=None, maxiter=1000, **kw):
        return train.optimize(self, method=method, tol=tol, callback=callback, maxiter=maxiter, **kw)

    # ------------------------------------------------------------------ predictions
    def predict_f(self, Xnew):
        """GPflow SGPR.predict_f (separation.py:306) -> mean [N*,1], var [N*,1]."""
        hyp, _ = self._hyp()
        m, v = self._engine().predict_f(_dev(np.asarray(Xnew).reshape(1, -1)), _dev(hyp), self._noise())
        return m[0].cpu().numpy().reshape(-1, 1), v[0].cpu().numpy().reshape(-1, 1)

    def predict_f_full_cov(self, Xnew):
        """GPflow Model.predict_f_full_cov -> mean [N*,1], cov [N*,N*,1]."""
        hyp, _ = self._hyp()
        m, v = self._engine().predict_f(_dev(np.asarray(Xnew).reshape(1, -1)), _dev(hyp), self._noise(), full_cov=True)
        return m[0].cpu().numpy().reshape(-1, 1), v[0].cpu().numpy()[:, :, None]

    def build_predict_source(self, Xnew, full_cov=False):
        """sgpr_ss.py:73-106: per-source dense GP posterior -> (list of P means [N*,1], list of P vars [N*,1], or
        [N*,N*,1] covariances when full_cov)."""
        hyp, _ = self._hyp()
        m, v = self._engine().predict_s(_dev(np.asarray(Xnew).reshape(1, -1)), _dev(hyp), self._noise(),
                                        full_cov=full_cov)
        P = m.shape[1]
        shape_v = (lambda a: a[:, :, None]) if full_cov else (lambda a: a.reshape(-1, 1))
        return ([m[0, i].cpu().numpy().reshape(-1, 1) for i in range(P)],
                [shape_v(v[0, i].cpu().numpy()) for i in range(P)])

    def predict_s(self, Xnew):
        """sgpr_ss.py:108-114."""
        return self.build_predict_source(Xnew)

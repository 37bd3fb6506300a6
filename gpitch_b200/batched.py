"""Window-batched engines: one ELBO (+ gradient) evaluation for W independent audio windows per call.

The reference fits windows one at a time in a Python loop (gpitch/separation.py:289-313,
gpitch/transcription.py:275-288); every window is an independent GP, so the window index is the batch dimension
of every kernel here and the sharding dimension across GPUs (gpitch_b200/distributed.py).

All tensors are torch fp64 CUDA tensors of CONSTRAINED parameter values, window-major:
    x, y [W, N];  noise [W]
Pdgp:   za [W,P,Ma], zc [W,P,Mc];  act_hyp [W,P,2] = (variance, lengthscale) of the Matern32 activation kernels;
        com_hyp [W,P,2+2Q] = (variance, lengthscale, energy[Q], frequency[Q]) of the MercerMatern12sm component
        kernels;  q_mu_* [W,P,M];  q_sqrt_* [W,P,M,M] (lower triangle used).
SGPR:   z [W,M];  hyp [W,P,2+2Q] (sum of P pitch kernels).
"""
import math
import os
import torch

from . import _lib as L
from .functions import (KernelMatrix, KernelPairOnGrid, SVGPConditional, SVGPConditionalG, SVGPConditionalHA, SGPRBound, VarExp, GaussKLWhite, Unwhiten,
                        cholesky_cond_estimate)

JITTER = 1e-6   # gpflow.settings.numerics.jitter_level


def _leaf(t):
    return t.detach().contiguous().requires_grad_(True)


class _nvtx(object):
    """NVTX range around a stage of the evaluation (chunk / latent-GP group / likelihood / backward), so that an
    nsys / ncu timeline of the batched step reads in the reference's terms.  A few hundred ns per range."""
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()


def _balanced_chunk(W, cw_max):
    """Equal-sized window chunks within the workspace budget: 256 windows at <= 17 per chunk -> 16 chunks of 16, not
    15 of 17 plus a single-window chunk whose kernels run at a fraction of the batch efficiency."""
    if W <= 0:
        return 1
    n = -(-W // cw_max)
    return -(-W // n)


def grid_lags(x, z):
    """Lag structure of inducing points that lie on their window's sample grid.  x [W, N] (uniform grids), z [R, M] with
    R = W * k rows, window-major.  Returns (iz int32 [R, M], delta [W], nlag) such that z[r, m] == x[r // k, iz[r, m]]
    exactly, or None when some point is off the grid or a grid is not uniform to a few ulps of its time stamps (then the
    general gradient kernel runs).  One device->host sync; call once per data set, not per evaluation."""
    W, N = x.shape
    R, M = z.shape
    if N < 2 or R % W:
        return None
    k = R // W
    x0 = x[:, :1]
    delta = ((x[:, -1] - x[:, 0]) / (N - 1)).contiguous()
    if not bool((delta > 0).all()):
        return None
    grid_err = (x - (x0 + torch.arange(N, dtype=x.dtype, device=x.device)[None, :] * delta[:, None])).abs().amax()
    tol = 8.0 * torch.finfo(x.dtype).eps * x.abs().amax()
    rows = torch.arange(R, device=x.device) // k
    iz = torch.round((z - x0[rows]) / delta[rows, None]).to(torch.int64)
    # pad points of ragged inducing sets (init_models.pad_inducing: >= 1e3 s outside the window, where the kernel is 0 to
    # 1e-44 and below) get an index no lag can reach: the histogram pass skips their rows, as their true weight is nil
    far = (z > x[rows, -1:] + 100.0) | (z < x0[rows] - 100.0)
    inside = (((iz >= 0) & (iz < N)) | far).all()
    on_grid = ((x.reshape(-1)[rows[:, None] * N + iz.clamp(0, N - 1)] == z) | far).all()
    if not bool(inside & on_grid & (grid_err <= tol)):
        return None
    top = int(torch.where(far, torch.zeros_like(iz), iz).max())
    iz = torch.where(far, torch.full_like(iz, -2 * N), iz)
    return iz.to(torch.int32).contiguous(), delta, int(N + top)


def _lag_slice(lag, sl, k):
    """Rows of a grid_lags() result that belong to the windows of slice `sl` (k point rows per window)."""
    if lag is None:
        return None
    iz, delta, nlag = lag
    return iz[sl.start * k:sl.stop * k], delta[sl], nlag


def mercer_kdiag(hyp):
    """MercerMatern12sm.Kdiag = variance * reduce(add, energy) (matern12_spectral_mixture.py:119-121); hyp [..., 2+2Q]."""
    Q = (hyp.shape[-1] - 2) // 2
    return hyp[..., 0] * hyp[..., 2:2 + Q].sum(-1)


class BatchedPdgp(object):
    """Pdgp.build_likelihood (gpitch/pdgp.py:133-170) for W windows x P pitches at once (whitened or not; optionally
    with gradients w.r.t. the inducing inputs, `train_z`)."""

    GFORM_COND_MAX = 1e4    # G-form rounding error ~ 6e-17 * cond(Kmm): 1e4 keeps it below 1e-12
    # formulation for groups the G-form is not certified for: 'ha' (3 M^2 N products) or 'tri' (4, GPflow's own order)
    STABLE_FORMS = {'ha': SVGPConditionalHA, 'tri': SVGPConditional}
    stable_form = 'ha'

    def __init__(self, x, y, za, zc, nlin='logistic', mode='reference', kind_com='mercer_m12', jitter=JITTER,
                 workspace_gb=24.0, gform='auto', whiten=True, train_z=False):
        self.x, self.y = x.contiguous(), y.contiguous()
        self.za, self.zc = za.contiguous(), zc.contiguous()
        self.W, self.N = x.shape
        self.P = za.shape[1]
        self.nlin, self.mode, self.kind_com, self.jitter = nlin, mode, kind_com, jitter
        self.workspace_gb = workspace_gb
        self.whiten = whiten
        self.train_z = train_z      # also return d ELBO / d za, d zc (Pdgp.za / zc left trainable, pdgp.py:80-85)
        self.two_streams = True
        self.two_streams_in_graph = True
        self.last_info = None
        # conditional() formulation per latent-GP group: True = G-form (2 M^2 N products), False = triangular form
        # (4 products, backward-stable for jitter-dominated Kmm), 'auto' = certified per group from the Cholesky
        # factors of the first evaluation (see functions.SVGPConditionalG); call reset_gform() after large
        # hyper-parameter moves.
        self.gform = {'act': gform, 'com': gform}
        self._gform_auto = {'act': gform == 'auto', 'com': gform == 'auto'}
        self._gform_age, self._gform_choice = {}, {}          # per (group, first window of the chunk)
        # component inducing points on the sample grid -> lag-histogram hyper-gradient (csrc/grad_lag.cu); 'auto' detects it
        # from the data on first use (one sync), False pins the general kernel
        self.lag_grad = False if os.environ.get('GPX_LAG_GRAD', '1') == '0' else 'auto'      # (env: A/B experiments)
        self._lag = None

    def set_data(self, x=None, y=None, za=None, zc=None):
        """Swap the windows' data in place (same shapes) without invalidating captured CUDA graphs."""
        for dst, src in ((self.x, x), (self.y, y), (self.za, za), (self.zc, zc)):
            if src is not None:
                dst.copy_(src)
        if self._lag not in (None, False) and (x is not None or zc is not None):
            new = grid_lags(self.x, self.zc.reshape(self.W * self.P, -1))
            if new is None or new[2] > self._lag[2]:
                raise ValueError('set_data: the new window does not have the grid structure the captured evaluation '
                                 'uses (inducing points off the sample grid); build a new engine')
            self._lag[0].copy_(new[0])          # in place: captured CUDA graphs keep reading these buffers
            self._lag[1].copy_(new[1])

    def _lag_info(self):
        """grid_lags() of the component group, detected once (never during graph capture, never while zc is trained)."""
        if self._lag is None:
            if self.lag_grad is False or self.train_z or self.kind_com != 'mercer_m12' or torch.cuda.is_current_stream_capturing():
                return None
            lag = grid_lags(self.x, self.zc.reshape(self.W * self.P, -1))
            # a window swap (set_data) may bring inducing points further right on the grid: size the histograms for any
            self._lag = (lag[0], lag[1], 2 * self.N) if lag else False
        return self._lag or None

    GFORM_RECHECK = 64      # evaluations between two re-certifications of an automatically chosen formulation

    def reset_gform(self, value='auto'):
        self.gform = {'act': value, 'com': value}
        self._gform_auto = {'act': value == 'auto', 'com': value == 'auto'}
        self._gform_age, self._gform_choice = {}, {}

    def _certify(self, group, chunk, Kmm):
        """G-form certificate of one chunk's group from the Cholesky factors of its Kmm (one device->host sync).
        Returns True when the stored choice changed.  `self.gform[group]` reports the conjunction over the chunks."""
        with torch.no_grad():
            est = float(cholesky_cond_estimate(Kmm.detach()).max())
        new = bool(est <= self.GFORM_COND_MAX)
        old = self._gform_choice.get((group, chunk))
        self._gform_choice[(group, chunk)] = new
        self._gform_age[(group, chunk)] = 0
        self.gform[group] = all(v for (g_, _), v in self._gform_choice.items() if g_ == group)
        return old is not None and old != new

    def _use_gform(self, group, Kmm, chunk=0):
        """Formulation of conditional() for a latent-GP group of one window chunk.  In 'auto' mode EVERY chunk is certified
        from the Cholesky factors of its own Kmm (per-window hyper-parameters differ between chunks) on its first evaluation
        and re-certified every GFORM_RECHECK evaluations, because hyper-parameters move during optimisation (a longer
        lengthscale raises cond(Kmm)); never during CUDA-graph capture -- graph owners call recertify_gform() instead."""
        if not self._gform_auto[group]:
            g = self.gform[group]
            return False if g == 'auto' else bool(g)
        if not self.whiten:
            # unwhitened models keep the stable form: their gradients (d/dz in particular) are differences of terms ~1e4
            # times their size, and the certificate is only a LOWER bound of cond(Kmm) -- a window certified on its own
            # showed 7e-7 on d ELBO / d zc in the G-form (tests: inducing_input_gradients[auto-False])
            return False
        key = (group, chunk)
        if not torch.cuda.is_current_stream_capturing():
            self._gform_age[key] = self._gform_age.get(key, 0) + 1
            if key not in self._gform_choice or self._gform_age[key] > self.GFORM_RECHECK:
                self._certify(group, chunk, Kmm)
        return bool(self._gform_choice.get(key, False))          # not certified yet (first call under capture): stable form

    def recertify_gform(self, act_hyp, com_hyp):
        """Eager re-certification of every chunk for the given hyper-parameters ([W, P, .]); True when a choice changed
        (the owner of a captured CUDA graph re-captures then -- the formulation is baked into the graph)."""
        changed = False
        P, cw = self.P, self.chunk_windows()
        for w0 in range(0, self.W, cw):
            sl = slice(w0, min(self.W, w0 + cw))
            Wc = sl.stop - sl.start
            for group, kind, hyp, z in (('act', 'matern32', act_hyp[sl].reshape(Wc * P, 1, 2), self.za[sl]),
                                        ('com', self.kind_com, com_hyp[sl].reshape(Wc * P, 1, -1), self.zc[sl])):
                if not self._gform_auto[group] or not self.whiten:
                    continue
                zz = z.reshape(Wc * P, -1)
                with torch.no_grad():
                    Kmm = KernelMatrix.apply(hyp.contiguous(), zz, zz, kind, self.mode, self.jitter, False)
                changed |= self._certify(group, w0, Kmm)
        return changed

    def chunk_windows(self):
        Ma, Mc = self.za.shape[2], self.zc.shape[2]
        M = max(Ma, Mc)
        per_gp = 8.0 * (5 * M * self.N + 14 * M * M)
        per_win = per_gp * 2 * self.P
        return _balanced_chunk(self.W, max(1, min(self.W, int(self.workspace_gb * 2 ** 30 / per_win))))

    def _group(self, kind, hyp, z, x, q_mu, q_sqrt, need_ef, group='com', lag=None, chunk=0):
        """One homogeneous group of Wc*P latent GPs -> fmean, fvar [Wc*P, N], kl [Wc*P], info."""
        Kmm = KernelMatrix.apply(hyp, z, z, kind, self.mode, self.jitter, need_ef)
        kdiag = hyp[:, 0, 0] if kind == 'matern32' else mercer_kdiag(hyp[:, 0, :])
        cond_fn = SVGPConditionalG if self._use_gform(group, Kmm, chunk) else self.STABLE_FORMS[self.stable_form]
        pre = (None, None, None)
        if not self.whiten:      # pdgp.py:122-129: evaluate the whitened model at (L^-1 q_mu, L^-1 Lq)
            q_mu, q_sqrt, *pre = Unwhiten.apply(q_mu, q_sqrt, Kmm)          # + the factor of Kmm, reused below
        if cond_fn is SVGPConditional:      # GPflow's operation order on a materialised Kmn / Kbar_mn
            Kmn = KernelMatrix.apply(hyp, z, x, kind, self.mode, 0.0, need_ef)
            fmean, fvar, info = cond_fn.apply(Kmn, Kmm, kdiag, q_mu, q_sqrt, *pre)
            kl = GaussKLWhite.apply(q_mu, q_sqrt)
        else:                               # fused stages build Kmn themselves, never write its adjoint, and own the KL term
            fmean, fvar, kl, info = cond_fn.apply(hyp, z, x, Kmm, kdiag, q_mu, q_sqrt, kind, self.mode, need_ef, *pre, lag)
        return fmean, fvar, kl, info

    NAMES = ('act_hyp', 'com_hyp', 'q_mu_act', 'q_sqrt_act', 'q_mu_com', 'q_sqrt_com', 'noise')

    def _elbo_chunk(self, sl, params, need_grad, need_ef, scale):
        """ELBO (+ gradients) of the windows in slice `sl`; params: dict of chunk tensors [Wc, ...]."""
        P, N = self.P, self.N
        Wc = params['noise'].shape[0]
        Ma, Mc = self.za.shape[2], self.zc.shape[2]
        with torch.set_grad_enabled(need_grad):
            leaf = {k: (_leaf(v) if need_grad else v.contiguous()) for k, v in params.items()}
            xa = self.x[sl]
            za, zc = self.za[sl].reshape(Wc * P, Ma), self.zc[sl].reshape(Wc * P, Mc)
            if need_grad and self.train_z:          # trainable inducing inputs (pdgp.py:80-85)
                za, zc = _leaf(za), _leaf(zc)
            # The activation and the component group are independent until the likelihood: they run on two side
            # streams so that the small-grid kernels of one (Cholesky panels, diagonal blocks, M x M x M products,
            # launch tails) overlap the other's large GEMMs.  autograd replays each group's backward on its stream.
            main = torch.cuda.current_stream()
            # (also while a CUDA graph is being captured: the fork / join below becomes two parallel branches of the graph,
            # which is what makes a single-window evaluation -- two equally long dependency chains -- ~1.7x faster)
            capturing = torch.cuda.is_current_stream_capturing()
            use_streams = self.two_streams and (self.two_streams_in_graph or not capturing)
            if use_streams:
                if not hasattr(self, '_s_grp'):
                    self._s_grp = (torch.cuda.Stream(), torch.cuda.Stream())
                s_a, s_c = self._s_grp
                s_a.wait_stream(main)
                s_c.wait_stream(main)
            else:
                s_a = s_c = main
            with torch.cuda.stream(s_a), _nvtx('pdgp.conditional[activations]'):
                fm_a, fv_a, kl_a, info_a = self._group('matern32', leaf['act_hyp'].reshape(Wc * P, 1, 2),
                                                       za, xa,
                                                       leaf['q_mu_act'].reshape(Wc * P, Ma),
                                                       leaf['q_sqrt_act'].reshape(Wc * P, Ma, Ma), False, 'act', chunk=sl.start or 0)
            with torch.cuda.stream(s_c), _nvtx('pdgp.conditional[components]'):
                fm_c, fv_c, kl_c, info_c = self._group(self.kind_com, leaf['com_hyp'].reshape(Wc * P, 1, -1),
                                                       zc, xa,
                                                       leaf['q_mu_com'].reshape(Wc * P, Mc),
                                                       leaf['q_sqrt_com'].reshape(Wc * P, Mc, Mc), need_ef,
                                                       lag=_lag_slice(self._lag_info(), sl, P) if need_grad else None, chunk=sl.start or 0)
            if use_streams:
                main.wait_stream(s_a)
                main.wait_stream(s_c)
                if not capturing:         # (a captured graph owns its memory pool: no cross-stream reuse to guard against)
                    for t in (fm_a, fv_a, kl_a, info_a, fm_c, fv_c, kl_c, info_c):
                        t.record_stream(main)
            Fmu = torch.cat([fm_a.view(Wc, P, N), fm_c.view(Wc, P, N)], 1)
            Fvar = torch.cat([fv_a.view(Wc, P, N), fv_c.view(Wc, P, N)], 1)
            with _nvtx('pdgp.variational_expectations'):
                ve = VarExp.apply(Fmu, Fvar, self.y[sl], leaf['noise'], self.nlin)
            kl = kl_a.view(Wc, P).sum(1) + kl_c.view(Wc, P).sum(1)
            elbo = ve * scale - kl
            g = None
            if need_grad:
                with _nvtx('pdgp.backward'):
                    elbo.sum().backward()
                g = {k: leaf[k].grad for k in self.NAMES}
                if self.train_z:
                    g['za'], g['zc'] = za.grad.view(Wc, P, Ma), zc.grad.view(Wc, P, Mc)
            if use_streams:            # everything the side streams touched (incl. the leaves) is done before reuse
                main.wait_stream(s_a)
                main.wait_stream(s_c)
        info = torch.stack([info_a.view(Wc, P), info_c.view(Wc, P)], 1)
        return elbo.detach(), g, info

    def elbo(self, act_hyp, com_hyp, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, noise, need_grad=True,
             need_ef=True, num_data=None):
        """Device-resident evaluation.  Returns elbo [W] and (if need_grad) a dict of gradients w.r.t. every
        argument (same shapes).  self.last_info [W, 2, P] holds the LAPACK-style status of every Cholesky."""
        W, N = self.W, self.N
        scale = 1.0 if num_data is None else float(num_data) / float(N)
        out = torch.empty(W, dtype=torch.float64, device=self.x.device)
        full = dict(zip(self.NAMES, (act_hyp, com_hyp, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com, noise)))
        grads = {k: torch.empty_like(v) for k, v in full.items()} if need_grad else None
        if need_grad and self.train_z:
            grads['za'], grads['zc'] = torch.empty_like(self.za), torch.empty_like(self.zc)
        infos = []
        cw = self.chunk_windows()
        for w0 in range(0, W, cw):
            sl = slice(w0, min(W, w0 + cw))
            e, g, info = self._elbo_chunk(sl, {k: v[sl] for k, v in full.items()}, need_grad, need_ef, scale)
            out[sl] = e
            if need_grad:
                for k in g:
                    grads[k][sl] = g[k]
            infos.append(info)
        self.last_info = torch.cat(infos, 0) if infos else torch.zeros((0, 2, self.P), dtype=torch.int32, device=self.x.device)
        return out, grads

    def elbo_host(self, params_host, elbo_host, grads_host, need_ef=True, num_data=None):
        """End-to-end evaluation from HOST buffers (what an optimiser driving the model from the host sees, like
        GPflow's Model._objective(x_free) -> (f, grad)): `params_host` / `grads_host` are dicts of pinned CPU
        tensors shaped like elbo()'s arguments (q_sqrt_* optionally as packed lower triangles [W, P, M (M + 1) / 2], in
        which case their gradients come back packed too), `elbo_host` a pinned [W] tensor.  Per window chunk the parameters
        are copied host->device on a copy stream, evaluated on the compute stream, and the gradients are copied
        device->host on a third stream, so the PCIe traffic overlaps the kernels of neighbouring chunks."""
        W, N = self.W, self.N
        scale = 1.0 if num_data is None else float(num_data) / float(N)
        dev = self.x.device
        main = torch.cuda.current_stream()
        if not hasattr(self, '_s_in'):
            self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_in, s_out = self._s_in, self._s_out
        cw = self.chunk_windows()
        chunks = [slice(w0, min(W, w0 + cw)) for w0 in range(0, W, cw)]
        # The first chunk's H2D copy and the last chunk's D2H copy cannot hide behind kernels: a short first and last chunk
        # (1/8 of a regular one) shrinks that exposed head and tail (0.5 GB each way per 32-window chunk of C3).
        h = cw // 8
        if h >= 1 and len(chunks) >= 2 and chunks[-1].stop - chunks[-1].start > h:
            first, last = chunks[0], chunks[-1]
            chunks = ([slice(first.start, first.start + h), slice(first.start + h, first.stop)] + chunks[1:-1] +
                      [slice(last.start, last.stop - h), slice(last.stop - h, last.stop)])
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        staged, done_ev, keep, infos = {}, {}, [], []
        # q_sqrt_* given as [W, P, M (M + 1) / 2]: packed lower triangles in, packed dLq out (half the PCIe bytes; only
        # tril(q_sqrt) is ever read and the gradient of the strict upper triangle is identically zero)
        packed = [k for k in ('q_sqrt_act', 'q_sqrt_com') if params_host[k].dim() == 3]

        def stage(i):
            with torch.cuda.stream(s_in):
                if i - 2 in done_ev:                       # double buffering: reuse after chunk i-2 was consumed
                    s_in.wait_event(done_ev[i - 2])
                d = {k: params_host[k][chunks[i]].to(dev, non_blocking=True) for k in self.NAMES}
                for k in packed:                           # packed lower triangles -> dense (zeros above the diagonal)
                    d[k] = L.tril_unpack(d[k], self.za.shape[2] if k == 'q_sqrt_act' else self.zc.shape[2])
                ev = torch.cuda.Event()
                ev.record(s_in)
            staged[i] = (d, ev)

        stage(0)
        if len(chunks) > 1:
            stage(1)
        for i, sl in enumerate(chunks):
            d, ev = staged.pop(i)
            main.wait_event(ev)
            for t in d.values():
                t.record_stream(main)
            e, g, info = self._elbo_chunk(sl, d, True, need_ef, scale)
            for k in packed:
                g[k] = L.tril_pack(g[k].contiguous())
            dev_done = torch.cuda.Event()
            dev_done.record(main)
            done_ev[i] = dev_done
            infos.append(info)
            with torch.cuda.stream(s_out):
                s_out.wait_event(dev_done)
                elbo_host[sl].copy_(e, non_blocking=True)
                for k in g:                                # 'za' / 'zc' too when the engine trains them
                    if k in grads_host:
                        grads_host[k][sl].copy_(g[k], non_blocking=True)
                    g[k].record_stream(s_out)
                e.record_stream(s_out)
            keep.append((e, g))
            if i + 2 < len(chunks):
                stage(i + 2)
        main.wait_stream(s_out)
        self.last_info = torch.cat(infos, 0)
        return elbo_host, grads_host

    @torch.no_grad()
    def predict(self, xnew, act_hyp, com_hyp, q_mu_act, q_sqrt_act, q_mu_com, q_sqrt_com):
        """Pdgp.predict_act_n_com (gpitch/pdgp.py:190-208): mean_a, var_a, mean_c, var_c [W,P,N*], mean_source."""
        W, P = self.W, self.P
        Ns = xnew.shape[1]
        Ma, Mc = self.za.shape[2], self.zc.shape[2]
        xnew = xnew.contiguous()
        fm_a, fv_a, _, _ = self._group('matern32', act_hyp.reshape(W * P, 1, 2).contiguous(), self.za.reshape(W * P, Ma),
                                       xnew, q_mu_act.reshape(W * P, Ma).contiguous(),
                                       q_sqrt_act.reshape(W * P, Ma, Ma).contiguous(), False, 'act', chunk='all')
        fm_c, fv_c, _, _ = self._group(self.kind_com, com_hyp.reshape(W * P, 1, -1).contiguous(),
                                       self.zc.reshape(W * P, Mc), xnew, q_mu_com.reshape(W * P, Mc).contiguous(),
                                       q_sqrt_com.reshape(W * P, Mc, Mc).contiguous(), False, chunk='all')
        ma, va = fm_a.view(W, P, Ns), fv_a.view(W, P, Ns)
        mc, vc = fm_c.view(W, P, Ns), fv_c.view(W, P, Ns)
        from .methods import nlin_torch
        return ma, va, mc, vc, nlin_torch(self.nlin)(ma) * mc


class GraphedEvaluation(object):
    """CUDA-graph replay of one ELBO(+gradient) evaluation for fixed shapes.  A single window (configs[0], [1]) is
    launch-bound: ~60 library calls and ~300 kernels of a few microseconds each per evaluation; captured once, an
    evaluation becomes one graph launch.  `fn(**static_inputs)` must be the engine call; inputs are copied into
    static buffers before every replay and outputs are returned as views of static tensors (clone to keep)."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.static_in = {k: v.clone() for k, v in example_inputs.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # also settles lazy initialisation (tables, gform choice)
                fn(**self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(**self.static_in)

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static_in[k].copy_(v)
        self.graph.replay()
        return self.static_out


class BatchedSGPR(object):
    """SGPRSS.build_likelihood / predict_f / predict_s (gpitch/sgpr_ss.py) for W windows at once; the kernel is
    the GPflow `Add` of P pitch kernels of one kind (gpitch/transcription.py:245, gpitch/separation.py:257)."""

    def __init__(self, x, y, z, kind='mercer_m12', mode='reference', reg=False, jitter=JITTER, workspace_gb=24.0):
        self.x, self.y, self.z = x.contiguous(), y.contiguous(), z.contiguous()
        self.W, self.N = x.shape
        self.M = z.shape[1]
        self.kind, self.mode, self.reg, self.jitter = kind, mode, reg, jitter
        self.workspace_gb = workspace_gb
        self.last_info = None
        # inducing points on the sample grid -> lag-histogram hyper-gradient (csrc/grad_lag.cu)
        self.lag_grad = False if os.environ.get('GPX_LAG_GRAD', '1') == '0' else 'auto'
        self._lag = None
        self.use_composite = os.environ.get('GPX_SGPR_COMPOSITE', '1') != '0'      # False: the autograd.Function path

    def set_data(self, x=None, y=None, z=None):
        """Swap the windows' data in place (same shapes): the DataHolder assignment of gpitch/separation.py:266-268
        without invalidating captured CUDA graphs."""
        for dst, src in ((self.x, x), (self.y, y), (self.z, z)):
            if src is not None:
                dst.copy_(src)
        if self._lag not in (None, False) and (x is not None or z is not None):
            new = grid_lags(self.x, self.z)
            if new is None or new[2] > self._lag[2]:
                raise ValueError('set_data: the new window does not have the grid structure the captured evaluation '
                                 'uses (inducing points off the sample grid); build a new engine')
            self._lag[0].copy_(new[0])          # in place: captured CUDA graphs keep reading these buffers
            self._lag[1].copy_(new[1])

    def _lag_info(self):
        if self._lag is None:
            if self.lag_grad is False or self.kind != 'mercer_m12' or torch.cuda.is_current_stream_capturing():
                return None
            lag = grid_lags(self.x, self.z)
            # a window swap (set_data) may bring inducing points further right on the grid: size the histograms for any
            self._lag = (lag[0], lag[1], 2 * self.N) if lag else False
        return self._lag or None

    def chunk_windows(self):
        per_win = 8.0 * (5 * self.M * self.N + 14 * self.M * self.M)
        return _balanced_chunk(self.W, max(1, min(self.W, int(self.workspace_gb * 2 ** 30 / per_win))))

    def _kdiag_sum(self, hyp, n):
        kd = hyp[:, :, 0] if self.kind == 'matern32' else mercer_kdiag(hyp)          # [Wc, P]
        return kd.sum(1) * n, kd.sum(1)

    def bound(self, hyp, noise, need_grad=True, need_ef=True):
        W, N = self.W, self.N
        out = torch.empty(W, dtype=torch.float64, device=self.x.device)
        grads = {'hyp': torch.empty_like(hyp), 'noise': torch.empty_like(noise)} if need_grad else None
        infos = []
        cw = self.chunk_windows()
        if self.use_composite:
            # the whole chunk in ONE C call (csrc/composite.cu: gpx_sgpr_bound) -- the same launch sequence as the
            # autograd.Function path below with its ~25 element-wise torch kernels fused into 8 small ones
            infos = []
            for w0 in range(0, W, cw):
                sl = slice(w0, min(W, w0 + cw))
                with _nvtx('sgprss.build_likelihood[chunk, composite]'):
                    b, dh, dn, info = L.sgpr_bound(self.kind, self.mode, self.x[sl], self.y[sl], self.z[sl], hyp[sl].contiguous(),
                                                   noise[sl].contiguous(), jitter=self.jitter, reg=self.reg,
                                                   lag=_lag_slice(self._lag_info(), sl, 1), need_grad=need_grad, need_ef=need_ef)
                out[sl] = b
                if need_grad:
                    grads['hyp'][sl] = dh
                    grads['noise'][sl] = dn
                infos.append(info.t())
            self.last_info = torch.cat(infos, 0) if infos else torch.zeros((0, 2), dtype=torch.int32, device=self.x.device)
            return out, grads
        for w0 in range(0, W, cw):
            sl = slice(w0, min(W, w0 + cw))
            with torch.set_grad_enabled(need_grad), _nvtx('sgprss.build_likelihood[chunk]'):
                h = _leaf(hyp[sl]) if need_grad else hyp[sl].contiguous()
                nv = _leaf(noise[sl]) if need_grad else noise[sl].contiguous()
                lag = _lag_slice(self._lag_info(), sl, 1)
                if lag is not None:         # inducing points on the sample grid: Kuu gathered from Kuf, one gradient pass
                    Kuf, Kuu = KernelPairOnGrid.apply(h, self.z[sl], self.x[sl], self.kind, self.mode, self.jitter, need_ef, lag)
                else:
                    Kuf = KernelMatrix.apply(h, self.z[sl], self.x[sl], self.kind, self.mode, 0.0, need_ef)
                    Kuu = KernelMatrix.apply(h, self.z[sl], self.z[sl], self.kind, self.mode, self.jitter, need_ef)
                skd, _ = self._kdiag_sum(h, N)
                bound, info = SGPRBound.apply(Kuf, Kuu, skd, self.y[sl], nv)
                if self.reg:                       # -1000 * sum_p |variance_p|   (sgpr_ss.py:64-68)
                    bound = bound - 1000.0 * h[:, :, 0].abs().sum(1)
                if need_grad:
                    bound.sum().backward()
                    grads['hyp'][sl] = h.grad
                    grads['noise'][sl] = nv.grad
            out[sl] = bound.detach()
            infos.append(info)
            del Kuf, Kuu, bound
        self.last_info = torch.cat(infos, 0) if infos else torch.zeros((0, 2), dtype=torch.int32, device=self.x.device)
        return out, grads

    @torch.no_grad()
    def _posterior(self, hyp, noise):
        z, x, y = self.z, self.x, self.y
        Kuf = KernelMatrix.apply(hyp, z, x, self.kind, self.mode, 0.0, False)
        Kuu = KernelMatrix.apply(hyp, z, z, self.kind, self.mode, self.jitter, False)
        inv_sigma = torch.rsqrt(noise).contiguous()
        Lm, Linv, _ = L.potrf_trinv(Kuu)
        A = L.gemm(Linv, Kuf, flags=L.GEMM_A_LOWER, alpha_vec=inv_sigma)
        B = L.gemm(A, A, flags=L.GEMM_TRANS_B | L.GEMM_C_LOWER | L.GEMM_C_MIRROR)
        B.diagonal(dim1=1, dim2=2).add_(1.0)
        LB, LBinv, _ = L.potrf_trinv(B)
        Aerr = L.rowdot(A, y)
        c = L.gemm(LBinv, Aerr.unsqueeze(2), flags=L.GEMM_A_LOWER).squeeze(2) * inv_sigma[:, None]
        return Linv, LBinv, c

    @torch.no_grad()
    def predict_f(self, xnew, hyp, noise, full_cov=False):
        """GPflow SGPR.build_predict(Xnew, full_cov) (separation.py:306) -> mean [W, N*], var [W, N*]
        (full_cov: [W, N*, N*] = K** + tmp2^T tmp2 - tmp1^T tmp1)."""
        hyp, noise, xnew = hyp.contiguous(), noise.contiguous(), xnew.contiguous()
        Linv, LBinv, c = self._posterior(hyp, noise)
        Kus = KernelMatrix.apply(hyp, self.z, xnew, self.kind, self.mode, 0.0, False)
        tmp1 = L.gemm(Linv, Kus, flags=L.GEMM_A_LOWER)
        tmp2 = L.gemm(LBinv, tmp1, flags=L.GEMM_A_LOWER)
        _, kd = self._kdiag_sum(hyp, xnew.shape[1])
        if full_cov:
            cov = KernelMatrix.apply(hyp, xnew, xnew, self.kind, self.mode, 0.0, False)
            L.gemm(tmp2, tmp2, out=cov, flags=L.GEMM_TRANS_A, beta=1.0)
            L.gemm(tmp1, tmp1, out=cov, flags=L.GEMM_TRANS_A, alpha=-1.0, beta=1.0)
            mean, _ = L.cond_colstats(tmp2, None, c.contiguous(), kd.contiguous())
            return mean, cov
        _, var = L.cond_colstats(tmp1, tmp2, c.contiguous(), kd.contiguous())
        mean, _ = L.cond_colstats(tmp2, None, c.contiguous(), kd.contiguous())
        return mean, var

    def _window_chunks(self, bytes_per_window):
        """Window slices such that one slice's temporaries stay inside the workspace budget."""
        cw = max(1, min(self.W, int(self.workspace_gb * 2 ** 30 / max(1.0, bytes_per_window))))
        return [slice(w0, min(self.W, w0 + cw)) for w0 in range(0, self.W, cw)]

    def _sub(self, sl):
        """Engine over the windows of slice `sl` (views of this engine's buffers)."""
        e = BatchedSGPR(self.x[sl], self.y[sl], self.z[sl], kind=self.kind, mode=self.mode, reg=self.reg,
                        jitter=self.jitter, workspace_gb=self.workspace_gb)
        return e

    @torch.no_grad()
    def predict_f_chunked(self, xnew, hyp, noise, full_cov=False):
        """predict_f in workspace-sized window slices (same results; bounded temporaries for whole tracks)."""
        Ns = xnew.shape[1]
        per_win = 8.0 * (4.0 * self.M * self.N + 4.0 * self.M * Ns + 14.0 * self.M * self.M + (Ns * Ns if full_cov else 0))
        ms, vs = [], []
        for sl in self._window_chunks(per_win):
            m, v = self._sub(sl).predict_f(xnew[sl], hyp[sl], noise[sl], full_cov=full_cov)
            ms.append(m); vs.append(v)
        return torch.cat(ms, 0), torch.cat(vs, 0)

    @torch.no_grad()
    def predict_s_chunked(self, xnew, hyp, noise, full_cov=False):
        """predict_s for many windows: the dense per-window posterior needs an N x N Cholesky per window (32 MB at
        N = 2001), so windows are processed in workspace-sized slices (a 4-minute track has 3838 of them)."""
        N, Ns = self.N, xnew.shape[1]
        per_win = 8.0 * (3.0 * N * N + 3.0 * N * Ns + (Ns * Ns * (1 + hyp.shape[1]) if full_cov else 0))
        ms, vs, infos = [], [], []
        for sl in self._window_chunks(per_win):
            sub = self._sub(sl)
            m, v = sub.predict_s(xnew[sl], hyp[sl], noise[sl], full_cov=full_cov)
            ms.append(m); vs.append(v); infos.append(sub.last_info)
        self.last_info = torch.cat(infos, 0)
        return torch.cat(ms, 0), torch.cat(vs, 0)

    @torch.no_grad()
    def predict_s(self, xnew, hyp, noise, full_cov=False):
        """SGPRSS.build_predict_source (sgpr_ss.py:73-106): dense GP per source.  Returns mean, var [W, P, N*]
        (full_cov: var [W, P, N*, N*]).
        NB var_i = Kdiag_sum(Xnew) - sum_n A_i^2 uses the SUM kernel's diagonal, exactly as the reference."""
        hyp, noise, xnew = hyp.contiguous(), noise.contiguous(), xnew.contiguous()
        W, P = hyp.shape[0], hyp.shape[1]
        Ns = xnew.shape[1]
        Kxx = KernelMatrix.apply(hyp, self.x, self.x, self.kind, self.mode, 0.0, False)
        Kxx.diagonal(dim1=1, dim2=2).add_(noise[:, None])
        _, Linv, info = L.potrf_trinv(Kxx)
        V = L.gemm(Linv, self.y.unsqueeze(2).contiguous(), flags=L.GEMM_A_LOWER).squeeze(2).contiguous()
        _, kd = self._kdiag_sum(hyp, Ns)
        means, vars_ = [], []
        for i in range(P):
            Kx = KernelMatrix.apply(hyp[:, i:i + 1, :].contiguous(), self.x, xnew, self.kind, self.mode, 0.0, False)
            A = L.gemm(Linv, Kx, flags=L.GEMM_A_LOWER)
            m, v = L.cond_colstats(A, None, V, kd.contiguous())
            if full_cov:       # sgpr_ss.py:99-100: K_sum(Xnew) - A^T A
                v = KernelMatrix.apply(hyp, xnew, xnew, self.kind, self.mode, 0.0, False)
                L.gemm(A, A, out=v, flags=L.GEMM_TRANS_A, alpha=-1.0, beta=1.0)
            means.append(m)
            vars_.append(v)
        self.last_info = info
        return torch.stack(means, 1), torch.stack(vars_, 1)

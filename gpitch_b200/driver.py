"""Window-batched optimiser drivers: what `SoSp.optimize` (gpitch/separation.py:279-313) and
`AMT.optimize` (gpitch/transcription.py:265-298) do one window at a time in a Python loop -- W independent
maximisations of the bound, advanced in lock-step so that every iteration is ONE batched ELBO+gradient evaluation.
The step logic (L-BFGS two-loop recursion, Armijo backtracking, Adam moments) stays on the host side of the ABI as
small torch tensor ops over [W, D] free-state matrices; windows never exchange information, so the result of each
window is what a per-window optimiser with the same rule would produce.

NB the reference's loops warm-start the (never reset) energies / frequencies of window i+1 from window i
(SURVEY.md 8(e) caveat); a lock-step batch starts every window from the same initial point instead.
"""
import torch

POS_LOWER = 1e-6   # GPflow transforms.positive: y = softplus(x) + 1e-6


def pos_forward(x):
    return torch.nn.functional.softplus(x) + POS_LOWER


def pos_backward(y):
    y = y - POS_LOWER
    return y + torch.log(-torch.expm1(-y))


class FreeState(object):
    """Packs the TRAINABLE entries of a dict of constrained [W, ...] tensors into a free matrix x [W, D] and back.
    `positive` names go through GPflow's positive transform, the rest are identity; `mask[name]` (bool, shape of one
    window's tensor) selects trainable entries (default: all)."""

    def __init__(self, params, positive=(), mask=None):
        self.names = list(params)
        self.template = {k: v.clone() for k, v in params.items()}
        self.positive = set(positive)
        self.idx = {}
        mask = mask or {}
        for k, v in params.items():
            m = mask.get(k)
            m = torch.ones(v.shape[1:], dtype=torch.bool, device=v.device) if m is None else m.to(v.device)
            self.idx[k] = m.reshape(-1).nonzero().squeeze(1)
        self.sizes = [int(self.idx[k].numel()) for k in self.names]
        self.D = sum(self.sizes)

    def pack(self, params):
        cols = []
        for k in self.names:
            v = params[k].reshape(params[k].shape[0], -1)[:, self.idx[k]]
            cols.append(pos_backward(v) if k in self.positive else v)
        return torch.cat(cols, 1)

    def unpack(self, x):
        """x [W, D] -> (constrained params dict, chain [W, D] = d constrained / d free)."""
        out, chain, off = {}, [], 0
        for k, n in zip(self.names, self.sizes):
            xf = x[:, off:off + n]
            off += n
            full = self.template[k].reshape(self.template[k].shape[0], -1).clone()
            if k in self.positive:
                full[:, self.idx[k]] = pos_forward(xf)
                chain.append(torch.sigmoid(xf))
            else:
                full[:, self.idx[k]] = xf
                chain.append(torch.ones_like(xf))
            out[k] = full.reshape(self.template[k].shape)
        return out, torch.cat(chain, 1)

    def pack_grad(self, grads, chain):
        cols = [grads[k].reshape(grads[k].shape[0], -1)[:, self.idx[k]] for k in self.names]
        return torch.cat(cols, 1) * chain


def _objective(fs, evaluate):
    def f(x):
        params, chain = fs.unpack(x)
        val, grads = evaluate(params)                     # val [W] (to MAXIMISE), grads dict
        g = fs.pack_grad(grads, chain)
        bad = ~torch.isfinite(val)
        val = torch.where(bad, torch.full_like(val, -float('inf')), val)
        g = torch.where(torch.isfinite(g), g, torch.zeros_like(g))       # GPflow zeroes non-finite gradient entries
        return -val, -g
    return f


def adam(fs, evaluate, x0, maxiter, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, callback=None):
    """tf.train.AdamOptimizer update rule (demos/scripts/demo-modgp.py:44-45) on every window at once."""
    f = _objective(fs, evaluate)
    x = x0.clone()
    m = torch.zeros_like(x)
    v = torch.zeros_like(x)
    hist = []
    for t in range(1, int(maxiter) + 1):
        val, g = f(x)
        hist.append(val.clone())
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        lr_t = lr * (1 - beta2 ** t) ** 0.5 / (1 - beta1 ** t)
        x = x - lr_t * m / (v.sqrt() + eps)
        if callback is not None:
            callback(t, x, val)
    return x, torch.stack(hist)


def lbfgs(fs, evaluate, x0, maxiter, history=10, c1=1e-4, max_ls=12, gtol=1e-5, callback=None):
    """Batched L-BFGS with Armijo backtracking: one two-loop recursion and one line search per window, all windows
    evaluated together each trial; windows whose projected gradient is below gtol are frozen.
    Every window keeps ITS OWN curvature history (a ring of `history` pairs with a per-window write pointer and validity
    mask): a window whose step produced no usable pair (s.y <= 0 happens routinely with Armijo-only line searches on
    non-convex bounds) simply skips the update -- its older pairs stay in place and keep shaping its direction, and its
    initial scaling comes from its own most recent good pair."""
    f = _objective(fs, evaluate)
    W, D = x0.shape
    H = int(history)
    x = x0.clone()
    val, g = f(x)
    S = torch.zeros((H, W, D), dtype=x.dtype, device=x.device)
    Y = torch.zeros_like(S)
    valid = torch.zeros((H, W), dtype=torch.bool, device=x.device)
    ptr = torch.zeros(W, dtype=torch.long, device=x.device)          # next slot to write, per window
    ar = torch.arange(W, device=x.device)
    hist = [val.clone()]
    active = torch.isfinite(val)
    for it in range(int(maxiter)):
        active = active & (g.abs().max(1).values > gtol)
        if not bool(active.any()):
            break
        # two-loop recursion, batched over windows; pair k = the k-th most recent pair OF EACH WINDOW
        q = g.clone()
        stack = []
        for k in range(H):
            idx = (ptr - 1 - k) % H
            s_k, y_k, v_k = S[idx, ar], Y[idx, ar], valid[idx, ar]
            rho = torch.where(v_k, 1.0 / (y_k * s_k).sum(1).clamp_min(1e-300), torch.zeros_like(val))
            a = rho * (s_k * q).sum(1)
            q = q - a[:, None] * y_k
            stack.append((a, rho, s_k, y_k))
        idx0 = (ptr - 1) % H
        s0, y0, v0 = S[idx0, ar], Y[idx0, ar], valid[idx0, ar]
        gamma = torch.where(v0, (s0 * y0).sum(1) / (y0 * y0).sum(1).clamp_min(1e-300), torch.ones_like(val))
        q = q * gamma[:, None]
        for a, rho, s_k, y_k in reversed(stack):
            b = rho * (y_k * q).sum(1)
            q = q + (a - b)[:, None] * s_k
        d = -q
        gd = (g * d).sum(1)
        bad_dir = gd >= 0                                   # not a descent direction: fall back to steepest descent
        d = torch.where(bad_dir[:, None], -g, d)
        gd = torch.where(bad_dir, -(g * g).sum(1), gd)
        t = torch.ones(W, dtype=x.dtype, device=x.device)
        no_pair = ~v0 | bad_dir                             # no curvature information yet: first step 1 / |g|_1
        t = torch.where(no_pair, (1.0 / g.abs().sum(1).clamp_min(1e-12)).clamp(max=1.0), t)
        t = torch.where(active, t, torch.zeros_like(t))
        accepted = ~active
        x_new, val_new, g_new = x.clone(), val.clone(), g.clone()
        for _ in range(max_ls):
            xt = x + t[:, None] * d
            vt, gt = f(xt)
            ok = (~accepted) & (vt <= val + c1 * t * gd)
            x_new = torch.where(ok[:, None], xt, x_new)
            val_new = torch.where(ok, vt, val_new)
            g_new = torch.where(ok[:, None], gt, g_new)
            accepted = accepted | ok
            if bool(accepted.all()):
                break
            t = torch.where(accepted, t, 0.5 * t)
        s = x_new - x
        y_ = g_new - g
        good = ((s * y_).sum(1) > 1e-12 * (y_ * y_).sum(1)) & accepted & active
        active = active & accepted                          # line search exhausted (rounding floor): freeze the window
        # windows with a usable pair write it into their own ring slot; the others leave their history untouched
        S[ptr, ar] = torch.where(good[:, None], s, S[ptr, ar])
        Y[ptr, ar] = torch.where(good[:, None], y_, Y[ptr, ar])
        valid[ptr, ar] = torch.where(good, torch.ones_like(good), valid[ptr, ar])
        ptr = torch.where(good, (ptr + 1) % H, ptr)
        x, val, g = x_new, val_new, g_new
        hist.append(val.clone())
        if callback is not None:
            callback(it + 1, x, val)
    return x, torch.stack(hist)


def fit_sgpr_windows(engine, hyp0, noise0, maxiter=100, method='lbfgs', train_cols=None, lr=0.01, need_ef=True):
    """Maximise the SGPRSS bound of every window of a BatchedSGPR engine.  hyp0 [W,P,2+2Q], noise0 [W] (constrained).
    train_cols: bool [2+2Q] selecting trainable hyper-parameter columns (default: all but the lengthscale, as
    init_kern_com(len_fixed=True) in gpitch/transcription.py:213-245).  Returns dict with the fitted `hyp`, `noise`,
    objective history [iters, W] and `matrix_var` [P, W] (gpitch/transcription.py:286-288)."""
    HS = hyp0.shape[2]
    if train_cols is None:
        train_cols = torch.ones(HS, dtype=torch.bool)
        train_cols[1] = False
    mask = {'hyp': train_cols[None, :].expand(hyp0.shape[1], HS).clone()}
    fs = FreeState({'hyp': hyp0, 'noise': noise0}, positive=('hyp', 'noise'), mask=mask)

    def evaluate(p):
        return engine.bound(p['hyp'].contiguous(), p['noise'].contiguous(), need_grad=True, need_ef=need_ef)
    x0 = fs.pack({'hyp': hyp0, 'noise': noise0})
    x, hist = (lbfgs if method == 'lbfgs' else adam)(fs, evaluate, x0, maxiter, **({} if method == 'lbfgs' else {'lr': lr}))
    out, _ = fs.unpack(x)
    return {'hyp': out['hyp'], 'noise': out['noise'], 'history': hist, 'matrix_var': out['hyp'][:, :, 0].t().contiguous()}


def fit_pdgp_windows(engine, params0, maxiter=100, lr=0.01, train_hyp=True):
    """Adam on every window of a BatchedPdgp engine (demo-modgp.py:44-45).  params0: dict with BatchedPdgp.NAMES."""
    names = engine.NAMES
    mask = {}
    if not train_hyp:
        for k in ('act_hyp', 'com_hyp'):
            mask[k] = torch.zeros(params0[k].shape[1:], dtype=torch.bool)
    fs = FreeState({k: params0[k] for k in names}, positive=('act_hyp', 'com_hyp', 'noise'), mask=mask)

    def evaluate(p):
        return engine.elbo(*[p[k].contiguous() for k in names], need_grad=True)
    x, hist = adam(fs, evaluate, fs.pack({k: params0[k] for k in names}), maxiter, lr=lr)
    out, _ = fs.unpack(x)
    out['history'] = hist
    return out


@torch.no_grad()
def predict_sources_merged(engine, hyp, noise, n, drop_last_of_224001=True):
    """The prediction tail of SoSp.optimize plus SoSp.predict_s (gpitch/separation.py:305-313, 341-379) for all windows
    at once: per-window mixture and source posteriors at the window's own inputs, then the Hann overlap-add of
    window_overlap.merged_mean / merged_variance ON THE DEVICE, so only the merged streams leave the GPU.
    engine: BatchedSGPR over the `windowed` test signal (W windows of ws samples, hop (ws-1)/2); hyp [W,P,2+2Q],
    noise [W]; n = length of the un-windowed signal.  Returns a dict with
      'mean_f', 'var_f' [W*ws]   (SoSp.predict_f(): plain concatenation, overlaps repeated -- as the reference does),
      'esource' = [[m_p, v_p] for every pitch p], each [n', 1] like SoSp.esource, n' = n - 1 for the reference's
      n == 224001 special case (separation.py:367-377)."""
    from . import window_overlap
    ws = engine.N
    mf, vf = engine.predict_f_chunked(engine.x, hyp, noise)
    ms, vs = engine.predict_s_chunked(engine.x, hyp, noise)
    P = ms.shape[1]
    esource = []
    for p in range(P):
        m = window_overlap.merged_mean_device(ms[:, p, :].contiguous(), ws, n)
        v = window_overlap.merged_variance_device(vs[:, p, :].contiguous(), ws, n)
        if drop_last_of_224001 and m.numel() == 224001:
            m, v = m[:-1], v[:-1]
        esource.append([m.reshape(-1, 1), v.reshape(-1, 1)])
    return {'mean_f': mf.reshape(-1), 'var_f': vf.reshape(-1), 'esource': esource}

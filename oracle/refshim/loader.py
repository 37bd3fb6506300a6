"""Load the UNMODIFIED reference modules from /root/reference under the tf/gpflow shims.

Python-2 semantics restored without touching the files:
  * ``a / b`` on two integers floors (AST pass: every Div becomes ``__py2div__(a, b)``);
  * builtin ``reduce``; implicit relative imports (``from likelihoods import MpdLik``);
  * ``scipy.signal.hann`` (removed from SciPy) -> ``scipy.signal.windows.hann``;
  * absent optional deps that the hot path never calls (matplotlib, soundfile, peakutils) -> stubs.
"""
import ast
import functools
import os
import sys
import types
import numpy as np

REF = os.environ.get('GPITCH_REFERENCE', '/root/reference')


def __py2div__(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) \
            and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


class _Py2Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id='__py2div__', ctx=ast.Load()), args=[node.left, node.right], keywords=[]),
                node)
        return node


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    from . import tf_shim, gpflow_shim
    sys.modules['tensorflow'] = tf_shim
    g = types.ModuleType('gpflow')
    for n in ('settings', 'param', 'kernels', 'likelihoods', 'quadrature', 'densities', 'mean_functions', 'model',
              'sgpr', 'minibatch', 'conditionals', 'kullback_leiblers'):
        sub = getattr(gpflow_shim, n)
        setattr(g, n, sub)
        if isinstance(sub, types.ModuleType):
            sys.modules['gpflow.' + n] = sub
    sys.modules['gpflow'] = g
    for n in ('peakutils', 'soundfile', 'matplotlib', 'h5py'):
        if n not in sys.modules:
            _stub(n)
    if 'matplotlib.pyplot' not in sys.modules:
        _stub('matplotlib.pyplot')
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, 'hann'):
        scipy.signal.hann = scipy.signal.windows.hann
    try:
        import scipy.fftpack  # noqa: F401
    except Exception:
        _stub('scipy.fftpack', fft=np.fft.fft)


def _load(modname, relpath, aliases=()):
    path = os.path.join(REF, relpath)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    tree = ast.fix_missing_locations(_Py2Div().visit(tree))
    mod = types.ModuleType(modname)
    mod.__file__ = path
    mod.__dict__['__py2div__'] = __py2div__
    mod.__dict__['reduce'] = functools.reduce
    sys.modules[modname] = mod
    for a in aliases:
        sys.modules[a] = mod
    exec(compile(tree, path, 'exec'), mod.__dict__)
    return mod


def load_reference():
    """Returns a namespace with the reference modules: methods, mm (matern12_spectral_mixture),
    likelihoods, sgpr_ss, pdgp, window_overlap, init_kernels."""
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree not found at %s (golden generation only runs in the build container)' % REF)
    install_shims()
    pkg = types.ModuleType('gpitch')
    pkg.__path__ = []
    sys.modules['gpitch'] = pkg
    ns = types.SimpleNamespace()
    ns.methods = _load('gpitch.methods', 'gpitch/methods.py', aliases=('methods',))
    pkg.methods = ns.methods
    for n in ('logistic', 'logistic_tf', 'softplus_tf', 'gaussfun_tf', 'midi2freq', 'find_ideal_f0'):
        setattr(pkg, n, getattr(ns.methods, n))
    ns.mm = _load('gpitch.matern12_spectral_mixture', 'gpitch/matern12_spectral_mixture.py',
                  aliases=('matern12_spectral_mixture',))
    ns.likelihoods = _load('gpitch.likelihoods', 'gpitch/likelihoods.py', aliases=('likelihoods',))
    pkg.likelihoods = ns.likelihoods
    ns.sgpr_ss = _load('gpitch.sgpr_ss', 'gpitch/sgpr_ss.py')
    ns.pdgp = _load('gpitch.pdgp', 'gpitch/pdgp.py')
    ns.window_overlap = _load('gpitch.window_overlap', 'gpitch/window_overlap.py')
    ns.init_kernels = _load('gpitch.init_kernels', 'gpitch/init_kernels.py')
    return ns

"""Single-window objective latency through the reference-named model classes (what scipy's L-BFGS-B calls):
SGPRSS config C1 (N=1600, M=200, P=3, Q=10) and Pdgp config C2 (N=4000, M=400, P=1), with and without CUDA-graph replay."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gpitch_b200 as gp
from gpitch_b200 import synthetic

def timeit(m, n=30):
    x = m.get_free_state()
    for _ in range(5):
        m._objective(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        f, g = m._objective(x)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, f

sp = synthetic.sgpr_problem(1, 1600, 200, 3, 10)
for graph in (False, True):
    kerns = gp.init_kernels.init_kern_com(3, [np.asarray(0.1)] * 3, [sp['hyp'][0, p, 2:12] for p in range(3)],
                                          [sp['hyp'][0, p, 12:] for p in range(3)], len_fixed=True)
    m = gp.SGPRSS(sp['x'][0].reshape(-1, 1), sp['y'][0].reshape(-1, 1), np.sum(kerns), sp['z'][0].reshape(-1, 1))
    m.use_cuda_graph = graph
    ms, f = timeit(m)
    print('SGPRSS C1  cuda_graph=%-5s  %.3f ms / objective (%.0f evals/s)   -bound = %.10g' % (graph, ms, 1e3 / ms, f))
pp = synthetic.pdgp_problem(1, 4000, 400, 1, 10)
for graph in (False, True):
    kc = gp.init_kernels.init_kern_com(1, [np.asarray(0.1)], [pp['com_hyp'][0, 0, 2:12]], [pp['com_hyp'][0, 0, 12:]], len_fixed=False)
    ka = gp.init_kernels.init_kern_act(1)
    z = [[pp['za'][0, 0].reshape(-1, 1)], [pp['zc'][0, 0].reshape(-1, 1)]]
    m = gp.Pdgp(pp['x'][0].reshape(-1, 1), pp['y'][0].reshape(-1, 1), z, [ka, kc])
    m.za.fixed = True; m.zc.fixed = True          # demos/scripts/demo-modgp.py:40-41
    m.use_cuda_graph = graph
    ms, f = timeit(m, 20)
    print('Pdgp   C2  cuda_graph=%-5s  %.3f ms / objective (%.0f evals/s)   -elbo = %.10g' % (graph, ms, 1e3 / ms, f))

import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from gpitch_b200 import synthetic, _lib
from gpitch_b200.batched import BatchedSGPR
T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).cuda()
pr = synthetic.sgpr_problem(1, 1600, 200, 3, 10)
eng = BatchedSGPR(T(pr['x']), T(pr['y']), T(pr['z']))
h, n = T(pr['hyp']), T(pr['noise'])
for _ in range(3):
    eng.bound(h, n)
torch.cuda.synchronize()

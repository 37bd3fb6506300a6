import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n ** 3 / best * 1e-9
# batched 400x400x4000
W = 512
a = torch.randn(W, 400, 400, dtype=torch.float64, device="cuda"); b = torch.randn(W, 400, 4000, dtype=torch.float64, device="cuda")
for _ in range(2): c = torch.bmm(a, b)
torch.cuda.synchronize(); best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = torch.bmm(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
res["bmm_512x400x400x4000_tflops"] = 2 * W * 400 * 400 * 4000 / best * 1e-9
L = torch.linalg.cholesky(a @ a.transpose(1, 2) + 400 * torch.eye(400, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize(); best = 1e9
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); x = torch.linalg.solve_triangular(L, b, upper=False); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
res["trsm_512x400x4000_ms"] = best
res["trsm_tflops"] = W * 400 * 400 * 4000 / best * 1e-9
print(json.dumps(res))

# cuBLAS DGEMM peak via torch.matmul(float64) -- library reference for the FP64 roofline denominator.
import json, torch, time
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = a @ b; torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2.0 * n**3 / best * 1e-9
    # sustained 3 s
    t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        c = a @ b; cnt += 1
        if cnt % 8 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained"] = 2.0 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))

"""cuBLAS FP64 GEMM / POTRF throughput via torch: the 'library' denominator for the dense path."""
import json, torch, time
dev = torch.device("cuda:0")
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: torch.matmul(a, b))
    print(json.dumps({"kernel": "cublas_dgemm", "n": n, "tflops": 2 * n**3 / ms * 1e-9, "ms": ms}))
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
ms = timeit(lambda: torch.linalg.cholesky(spd), n=3)
print(json.dumps({"kernel": "cusolver_dpotrf", "n": n, "tflops": n**3 / 3 / ms * 1e-9, "ms": ms}))
x = torch.empty(1 << 28, dtype=torch.float64, device=dev); y = torch.empty_like(x)
ms = timeit(lambda: y.copy_(x))
print(json.dumps({"kernel": "copy_f64", "gbs": 2 * x.numel() * 8 / ms * 1e-6, "ms": ms}))
ms = timeit(lambda: y.fill_(1.0))
print(json.dumps({"kernel": "fill_f64", "gbs": x.numel() * 8 / ms * 1e-6, "ms": ms}))

"""cuBLAS DGEMM cross-check for the FP64 roofline denominator (comparator only, never on the product path)."""
import json, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"bench": "cublas_dgemm", "n": n, "ms": best, "tflops": 2 * n**3 / best * 1e-9}))
# tall-skinny shapes of the hot path (K1: M x L @ L x H, K2: L x M @ M x H), reduced
L, M, H = 20000, 50000, 64
Y = torch.randn(M, L, dtype=torch.float64, device=dev)  # = Julia Y (L x M col-major) viewed row-major
B = torch.randn(L, H, dtype=torch.float64, device=dev)
A = torch.randn(M, H, dtype=torch.float64, device=dev)
for name, f in (("YtB", lambda: Y @ B), ("YA", lambda: Y.t() @ A)):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"bench": "cublas_" + name, "L": L, "M": M, "H": H, "ms": best, "tflops": 2.0 * L * M * H / best * 1e-9}))

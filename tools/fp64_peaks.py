"""cuBLAS DGEMM peak (torch.matmul fp64), timed with CUDA events; companion of tools/fp64_peaks.cu."""
import json, sys, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    c = a @ b
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained: back-to-back for ~3 s
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = max(3, int(3000 / best))
e0.record()
for _ in range(reps):
    c = a @ b
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / reps
print(json.dumps({"kind": "cublas_dgemm", "n": n, "ms_best": best, "tflops_burst": 2 * n**3 / best * 1e-9,
                  "ms_sustained": sus, "tflops_sustained": 2 * n**3 / sus * 1e-9}))

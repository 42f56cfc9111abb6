"""Measure the FP64 GEMM peak (cuBLAS DGEMM through torch) -- the DMMA roofline denominator."""
import json, torch
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda"); b = torch.randn(N, N, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"fp64_dgemm_tflops": 2 * N**3 / best / 1e9, "ms": best, "N": N, "gpu": torch.cuda.get_device_name(0)}))

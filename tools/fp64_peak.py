"""Measure the FP64 GEMM peak (cuBLAS DGEMM through torch) -- the DMMA roofline denominator -- with the SM clock sampled
while it runs (burst = best of 10 isolated calls, sustained = back-to-back for 3 s)."""
import json, subprocess, threading, time, torch
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda"); b = torch.randn(N, N, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
clk, stop = [], threading.Event()
def sample():
    while not stop.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
            clk.append(o)
        except Exception:
            pass
        stop.wait(0.2)
th = threading.Thread(target=sample, daemon=True); th.start()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); n = 0; t0 = time.time()
while time.time() - t0 < 3.0:
    c = a @ b; n += 1
    if n % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / n
stop.set(); th.join(timeout=2)
print(json.dumps({"fp64_dgemm_tflops": 2 * N**3 / best / 1e9, "ms": best, "fp64_dgemm_tflops_sustained": 2 * N**3 / sus / 1e9,
                  "ms_sustained": sus, "N": N, "gpu": torch.cuda.get_device_name(0),
                  "clock_samples_sm_mhz,max,power_w,reasons": clk}))

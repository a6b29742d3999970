"""cuBLAS DGEMM peak on this box: the FP64 roofline denominator."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = a @ b; torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); k = 0; t0 = time.time()
while time.time() - t0 < 4.0:
    for _ in range(5): c = a @ b
    k += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"fp64_dgemm_tflops_burst": burst, "fp64_dgemm_tflops_sustained": sus, "n": n}))

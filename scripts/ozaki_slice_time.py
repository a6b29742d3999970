"""Time of the digit slicing kernel alone (8192 x 8192 fp64 -> S int8 planes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import torch
from mcpilco_b200 import _ops as ops, _native as Nn
N = 8192
g = torch.Generator(device="cuda").manual_seed(1)
B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
L = ops._enter(B.device)
for S in (8, 7):
    planes = torch.empty(L.mcpilco_ozaki_plane_bytes(N, S), dtype=torch.uint8, device="cuda")
    pexp = torch.empty(N, dtype=torch.int32, device="cuda")
    run = lambda: Nn.check(L.mcpilco_ozaki_prepare(ops._ptr(B), N, B.stride(0), S, ops._ptr(planes), ops._ptr(pexp), ops._stream(B.device)))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("slice S=%d: %.3f ms  (%.0f GB/s of fp64 in + int8 out)" % (S, ms, (N * N * 8 + N * N * S) / ms * 1e-6), flush=True)
    print("  checksum", int(planes.to(torch.int64).sum()), int(pexp.sum()))

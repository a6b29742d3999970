"""Correctness ladder and timing of the hand-written tcgen05 digit-plane kernel (ozaki_mma_kernel) against an fp64 matmul.
usage: python scripts/ozaki_mma_debug.py [--big]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import torch
from mcpilco_b200 import _ops as ops

def check(M, N, S, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
    ref = A @ B.t()
    V, _, _ = ops.ozaki_matmul(A, B, S)
    torch.cuda.synchronize()
    scale = A.abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).t() * N
    err = float(((V - ref).abs() / scale).max())
    bad = int((((V - ref).abs() / scale) > 2.0 ** -45).sum())
    print("M=%d N=%d S=%d: max scaled err %.3e  bad %d / %d  nan %d" % (M, N, S, err, bad, V.numel(), int(torch.isnan(V).sum())), flush=True)
    if bad:
        idx = (((V - ref).abs() / scale) > 2.0 ** -45).nonzero()
        print("   first bad:", idx[:5].tolist(), "rows with errors:", sorted(set(idx[:, 0].tolist()))[:10], "cols:", sorted(set(idx[:, 1].tolist()))[:10], flush=True)
    return err

for M, N in ((128, 128), (256, 256), (256, 512), (300, 1000), (513, 1100), (1024, 2048)):
    for S in (8, 7, 2):
        check(M, N, S)
if "--big" in sys.argv:
    for N in (4096, 8192):
        M = 8192
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
        L = ops._enter(A.device)
        import ctypes as C
        from mcpilco_b200 import _native as Nn
        for S in (8, 7):
            planes = torch.empty(L.mcpilco_ozaki_plane_bytes(N, S), dtype=torch.uint8, device="cuda")
            pexp = torch.empty(N, dtype=torch.int32, device="cuda")
            Nn.check(L.mcpilco_ozaki_prepare(ops._ptr(B), N, B.stride(0), S, ops._ptr(planes), ops._ptr(pexp), ops._stream(A.device)))
            V = torch.empty(M, N, dtype=torch.float64, device="cuda")
            sb = L.mcpilco_ozaki_scratch_bytes(M, N, S)
            scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
            def run():
                Nn.check(L.mcpilco_ozaki_contract(ops._ptr(A), A.stride(0), M, N, S, ops._ptr(planes), ops._ptr(pexp), ops._ptr(V), N, ops._ptr(scratch), sb, ops._stream(A.device)))
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            nprod = S * (S + 1) // 2
            print("N=%d M=%d S=%d: %.3f ms per contraction (slice + mma) = %.1f TFLOP/s fp64-equivalent, int8 rate %.2f POPS" % (N, M, S, ms, 2.0 * M * N * N / ms * 1e-9, nprod * 2.0 * M * N * N / ms * 1e-12), flush=True)
            ref = A[:64] @ B.t()
            print("   err vs fp64 on 64 rows: %.3e" % float(((V[:64] - ref).abs() / (A[:64].abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).t() * N)).max()), flush=True)

"""One contraction through the hand-written tcgen05 digit-plane kernel (for ncu): python scripts/ozaki_mma_one.py M N S"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import torch
from mcpilco_b200 import _ops as ops
M, N, S = (int(a) for a in sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    V, _, _ = ops.ozaki_matmul(A, B, S)
torch.cuda.synchronize()
print("ok", float(V[0, 0]))

"""One posterior (K* rows + digit planes, INT8 contraction, reduce) at the bench size, for an ncu launch list.
usage: python scripts/ozaki_predict_one.py [S] [N] [M]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
from mcpilco_b200 import _ops as ops, _pack as P, workloads as W
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
M = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
dev = "cuda:0"
sc = W.cartpole_sweep(N)
g = sc["gps"][0]
spec = P.spec_from_dict({"D": 6, "log_ls": g["log_ls"], "lambda": 1.0, "mean": 0.0, "mpk": g["mpk"], "sigma_n": 0.1})
X = torch.tensor(sc["X"], device=dev); y = torch.tensor(sc["Y"][:, :1].copy(), device=dev)
alpha, Kinv = ops.gp_precompute(spec, X, y)
rs = np.random.RandomState(0)
Xs = torch.tensor(sc["X"][rs.choice(N, M)] + 0.05 * rs.randn(M, 6), device=dev)
gp = ops.FittedGp(spec, X, alpha, Kinv, ozaki_slices=S)
for _ in range(3):
    out = ops.gp_predict([gp], Xs, jac=True)
torch.cuda.synchronize()
print("ok", float(out[1].sum()))

"""One GP precompute at N (default 8192) — the workload for an ncu launch list of the factorisation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import torch
from mcpilco_b200 import _ops as ops, _pack as P, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sc = W.cartpole_sweep(n)
X = torch.tensor(sc["X"], device="cuda:0"); y = torch.tensor(sc["Y"][:, :1].copy(), device="cuda:0")
g = sc["gps"][0]
spec = P.spec_from_dict({"D": 6, "log_ls": g["log_ls"], "lambda": 1.0, "mean": 0.0, "mpk": g["mpk"], "sigma_n": 0.1})
alpha, Kinv = ops.gp_precompute(spec, X, y)
torch.cuda.synchronize()
print("ok", float(alpha.abs().max()))

import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
from mcpilco_b200 import _ops as ops, _pack as P, workloads as W
sc = W.cartpole_sweep(8192)
dev = "cuda:0"
X = torch.tensor(sc["X"], device=dev); y = torch.tensor(sc["Y"][:, :1].copy(), device=dev)
g = sc["gps"][0]
spec = P.spec_from_dict({"D": 6, "log_ls": g["log_ls"], "lambda": 1.0, "mean": 0.0, "mpk": g["mpk"], "sigma_n": 0.1})
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    alpha, Kinv = ops.gp_precompute(spec, X, y)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    gp = ops.FittedGp(spec, X, alpha, Kinv)
    mean, var = ops.gp_predict([gp], X)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("iter %d: precompute %.1f ms, predict(8192) %.1f ms" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    time.sleep(3 if it == 1 else 0)

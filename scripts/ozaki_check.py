"""Accuracy and speed of the opt-in INT8 (Ozaki) posterior contraction against the native FP64 path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
from mcpilco_b200 import _ops as ops, _pack as P, workloads as W
dev = "cuda:0"
for N, M in ((1000, 513), (2048, 2048), (8192, 8192)):
    sc = W.cartpole_sweep(N)
    g = sc["gps"][0]
    spec = P.spec_from_dict({"D": 6, "log_ls": g["log_ls"], "lambda": 1.0, "mean": 0.0, "mpk": g["mpk"], "sigma_n": 0.1})
    X = torch.tensor(sc["X"], device=dev); y = torch.tensor(sc["Y"][:, :1].copy(), device=dev)
    alpha, Kinv = ops.gp_precompute(spec, X, y)
    rs = np.random.RandomState(0)
    Xs = torch.tensor(sc["X"][rs.choice(N, M)] + 0.05 * rs.randn(M, 6), device=dev)
    ref = ops.FittedGp(spec, X, alpha, Kinv, ozaki_slices=0)
    m0, v0, jm0, jv0 = ops.gp_predict([ref], Xs, jac=True)
    def timeit(gp):
        ops.gp_predict([gp], Xs, jac=True); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            ops.gp_predict([gp], Xs, jac=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / 3
    t_nat = timeit(ref)
    print("N=%d M=%d native: %.2f ms (%.1f TFLOP/s)  var/k** median %.1e" % (N, M, 1e3 * t_nat, 2.0 * M * N * N / t_nat * 1e-12,
          float((v0[:, 0] / ops.gp_diag_covariance(spec, Xs)).median())), flush=True)
    for S in (8, 7):
        gp = ops.FittedGp(spec, X, alpha, Kinv, ozaki_slices=S)
        m1, v1, jm1, jv1 = ops.gp_predict([gp], Xs, jac=True)
        t = timeit(gp)
        rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())
        reln = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print("   ozaki S=%d: %.2f ms (%.1f TFLOP/s fp64-equivalent, x%.2f)  var rel err max %.2e median %.2e | mean %.1e | jvar normwise %.2e" % (
            S, 1e3 * t, 2.0 * M * N * N / t * 1e-12, t_nat / t, rel(v1, v0), float(((v1 - v0).abs() / v0.abs()).median()), reln(m1, m0), reln(jv1, jv0)), flush=True)

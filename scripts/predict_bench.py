"""Times the posterior (covariance tile + K* Kinv GEMM + reduce) at sweep sizes; prints GEMM-equivalent TFLOP/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
from mcpilco_b200 import _ops as ops, _pack as P

def main():
    dev = "cuda:0"
    rs = np.random.RandomState(0)
    for N, M in ((2048, 16384), (8192, 16384), (8192, 4096)):
        X = torch.tensor(rs.uniform(-2, 2, (N, 6)), dtype=torch.float64, device=dev)
        y = torch.tensor(rs.randn(N, 1), dtype=torch.float64, device=dev)
        spec = P.spec_from_dict({"D": 6, "log_ls": [2, 2, 2, 0.8, 1.5, 2.5], "lambda": 1.0, "mean": 0.0,
                                 "mpk": [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)], "sigma_n": 0.1})
        torch.cuda.synchronize(); t0 = time.time()
        alpha, Kinv = ops.gp_precompute(spec, X, y)
        torch.cuda.synchronize(); tp = time.time() - t0
        gp = ops.FittedGp(spec, X, alpha, Kinv)
        Xs = torch.tensor(rs.uniform(-2, 2, (M, 6)), dtype=torch.float64, device=dev)
        for jac in (False, True):
            ops.gp_predict([gp], Xs, jac=jac)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(3):
                ops.gp_predict([gp], Xs, jac=jac)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print("N=%d M=%d jac=%d: %.2f ms  %.2f TFLOP/s (2MN^2)  precompute %.1f ms" % (N, M, jac, ms, 2.0 * M * N * N / ms * 1e-9, tp * 1e3), flush=True)

main()

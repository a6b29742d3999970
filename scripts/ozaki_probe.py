"""FEASIBILITY PROBE for SURVEY.md §8 f4 (not a product path): fp64 GEMM V = K* K^-1 emulated with int8 tensor-core GEMMs
(Ozaki scheme I: S slices of 7 bits per operand, the S(S+1)/2 slice products with t + u < S accumulated in int32 per weight).
Here the int8 GEMMs are torch._int_mm (cuBLASLt) — a library stand-in used only to learn (a) how many slices the posterior variance
needs and (b) what an int8 tensor-core kernel would have to sustain to beat the native FP64 pipe.  A product version would be a
hand-written tcgen05 (kind::i8, TMEM accumulators) kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
from mcpilco_b200 import _ops as ops, _pack as P, workloads as W

dev = "cuda:0"
W_BITS = 7


def slices(A, S):
    """Row-scaled signed 7-bit digits: A = 2^e[:,None] * sum_t D_t 2^{-7(t+1)} (+ remainder < 2^{-7S})."""
    amax = A.abs().amax(1, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax)) + 1.0          # |A| 2^-e < 0.5
    r = A * torch.exp2(-e)
    out = []
    for _ in range(S):
        r = r * (2.0 ** W_BITS)
        d = torch.trunc(r)
        r = r - d
        out.append(d.to(torch.int8))
    return out, e


def ozaki_gemm(A, B, S, timing=None):
    """A [M,K] @ B[N,K]^T with int8 GEMMs."""
    As, eA = slices(A, S)
    Bs, eB = slices(B, S)
    BsT = [b.t().contiguous().t() for b in Bs]  # [K,N] column-major views as _int_mm likes
    M, N = A.shape[0], B.shape[0]
    V = torch.zeros(M, N, dtype=torch.float64, device=A.device)
    torch.cuda.synchronize(); t0 = time.perf_counter(); n_mm = 0
    for d in range(S):
        C = torch.zeros(M, N, dtype=torch.int32, device=A.device)
        for t in range(d + 1):
            C += torch._int_mm(As[t], Bs[d - t].t())
            n_mm += 1
        V += C.to(torch.float64) * (2.0 ** (-W_BITS * (d + 2)))
    V = V * torch.exp2(eA) * torch.exp2(eB).t()
    torch.cuda.synchronize()
    if timing is not None:
        timing.append((time.perf_counter() - t0, n_mm))
    return V


def main():
    N, M = 8192, 4096
    sc = W.cartpole_sweep(N)
    g = sc["gps"][0]
    spec = P.spec_from_dict({"D": 6, "log_ls": g["log_ls"], "lambda": 1.0, "mean": 0.0, "mpk": g["mpk"], "sigma_n": 0.1})
    X = torch.tensor(sc["X"], device=dev); y = torch.tensor(sc["Y"][:, :1].copy(), device=dev)
    alpha, Kinv = ops.gp_precompute(spec, X, y)
    rs = np.random.RandomState(0)
    Xs = torch.tensor(sc["X"][rs.choice(N, M)] + 0.05 * rs.randn(M, 6), device=dev)
    Ks = ops.gp_covariance(spec, Xs, X)
    Kinv = Kinv.contiguous()
    V64 = Ks @ Kinv
    kd = ops.gp_diag_covariance(spec, Xs)
    var64 = kd - (V64 * Ks).sum(1)
    # a higher-precision yardstick for the quadratic form: split Ks into hi/lo and use fp64 twice (error-free-ish in the A operand)
    print("N=%d M=%d  var/k** median %.2e  |Kinv|max %.2e" % (N, M, float((var64 / kd).median()), float(Kinv.abs().max())))
    mean_native, var_native = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv)], Xs)
    print("native CUDA path vs torch fp64 (cuBLAS): var rel diff median %.2e max %.2e" % (
        float(((var_native[:, 0] - var64).abs() / var64.abs()).median()), float(((var_native[:, 0] - var64).abs() / var64.abs()).max())))
    for S in (6, 7, 8, 9, 10):
        tm = []
        V = ozaki_gemm(Ks, Kinv, S, tm)
        var = kd - (V * Ks).sum(1)
        relV = float((V - V64).abs().max() / V64.abs().max())
        relvar = (var - var64).abs() / var64.abs()
        sec, n_mm = tm[0]
        print("S=%2d (%2d int8 GEMMs): max|dV|/max|V| %.2e   var rel diff median %.2e  max %.2e   wall %.1f ms -> %.1f TFLOP/s fp64-equivalent "
              "(int8 %.2f POP/s incl. int32 adds, slicing excluded)" % (S, n_mm, relV, float(relvar.median()), float(relvar.max()), 1e3 * sec,
                                                                       2.0 * M * N * N / sec * 1e-12, 2.0 * M * N * N * n_mm / sec * 1e-15), flush=True)
    # pure int8 GEMM rate of the library at this shape (upper bound for a hand-written tcgen05 kernel to aim at)
    a = torch.randint(-127, 127, (M, N), dtype=torch.int8, device=dev); b = torch.randint(-127, 127, (N, N), dtype=torch.int8, device=dev).t()
    torch._int_mm(a, b); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        torch._int_mm(a, b)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("torch._int_mm %dx%dx%d: %.3f ms = %.2f POP/s" % (M, N, N, 1e3 * dt, 2.0 * M * N * N / dt * 1e-15))

main()

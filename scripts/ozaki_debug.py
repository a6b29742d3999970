import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))
import numpy as np, torch
import ctypes as C
from mcpilco_b200 import _ops as ops, _native as Nn
torch.manual_seed(0)
for (M, N, S) in ((128, 128, 2), (128, 128, 7), (64, 300, 7), (256, 256, 8)):
    A = torch.randn(M, N, dtype=torch.float64, device="cuda"); B = torch.randn(N, N, dtype=torch.float64, device="cuda")
    V, planesB, expB = ops.ozaki_matmul(A, B, S)
    # slice A the same way (prepare() with reverse planes; undo the reversal)
    L = Nn.lib()
    Kp = (N + 127) // 128 * 128
    # emulate A planes on the host
    def digits(X):
        x = X.cpu().numpy(); R = x.shape[0]
        E = np.floor(np.log2(np.abs(x).max(1))).astype(int) + 3
        out = np.zeros((R, S, x.shape[1]), dtype=np.int64)
        for r in range(R):
            F = np.rint(np.ldexp(x[r], 8 * S - int(E[r]))).astype(np.int64)
            for t in range(S - 1, -1, -1):
                d = ((F + 128) & 255) - 128; F = (F - d) >> 8; out[r, t] = d
        return out, E
    dA, eA = digits(A); dB, eB = digits(B)
    Vemu = np.zeros((M, N)); per_w = []
    for w in range(S - 1, -1, -1):
        Cw = sum(dA[:, t, :] @ dB[:, w - t, :].T for t in range(w + 1))
        per_w.append((w, Cw))
        Vemu += Cw.astype(np.float64) * 256.0 ** -(w + 2)
    Vemu *= np.exp2(eA)[:, None] * np.exp2(eB)[None, :]
    ref = (A @ B.t()).cpu().numpy(); Vk = V.cpu().numpy()
    print("M=%d N=%d S=%d: emulation vs fp64 %.2e | kernel vs emulation %.2e | kernel vs fp64 %.2e" % (
        M, N, S, np.abs(Vemu - ref).max() / np.abs(ref).max(), np.abs(Vk - Vemu).max() / np.abs(ref).max(), np.abs(Vk - ref).max() / np.abs(ref).max()))
    # which weights are missing?  fit kernel V as sum_w c_w * term_w
    terms = np.stack([(Cw.astype(np.float64) * 256.0 ** -(w + 2) * np.exp2(eA)[:, None] * np.exp2(eB)[None, :]).ravel() for w, Cw in per_w], 1)
    coef = np.linalg.lstsq(terms, Vk.ravel(), rcond=None)[0]
    print("   least-squares weight of each C_w in the kernel output (w = %s): %s" % ([w for w, _ in per_w], np.round(coef, 4).tolist()))

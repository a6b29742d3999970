"""Host-side enqueue cost of one rollout forward/backward (the C call returns when everything is queued)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import native_helpers as nh
from mcpilco_b200 import workloads as W, _ops as ops
for N, M in ((2048, 256), (2048, 2048), (512, 256)):
    sc = W.cartpole_sweep(N); sc.update(M=M, H=60)
    gps = nh.native_fit(sc)
    plan, _ = nh.native_plan(sc, gps, need_grad=True, inject=False, seed=1)
    x0 = nh.G(sc["x0_mean"]).repeat(M, 1)
    for _ in range(2):
        plan.forward(x0); plan.backward(grad_cost=1.0)
    torch.cuda.synchronize()
    for prof in (False, True):
        ops.prof_enable(prof)
        t0 = time.perf_counter(); plan.forward(x0); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        plan.backward(grad_cost=1.0); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        print("N=%d M=%d prof=%d: forward enqueue %.2f ms (+%.2f ms to drain), backward enqueue %.2f ms (+%.2f)" % (N, M, prof, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
        if prof:
            ops.prof_read(); ops.prof_enable(False)

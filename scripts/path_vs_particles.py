"""Forward rollout time of the whole-horizon cluster kernel against the fused two-launch-per-step path as the particle count grows
(cart-pole shapes, H = 60).  usage: python scripts/path_vs_particles.py"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch


def run():
    import native_helpers as nh
    from mcpilco_b200 import workloads as W
    out = {}
    for key, N in (("c1", 300), ("c2", 300), ("c2", 120)):
        for M in (200, 400, 800, 1200, 2048):
            sc = W.real_shape(key, N=N, M=M, H=60, with_noise=False)
            sc["eps0"] = np.random.RandomState(1).randn(M, sc["Ds"])
            gps = nh.native_fit(sc)
            plan, _ = nh.native_plan(sc, gps, need_grad=True, inject=False, seed=1)
            x0 = nh.x0_of(sc)
            for _ in range(3):
                plan.forward(x0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                plan.forward(x0)
            e1.record(); torch.cuda.synchronize()
            out["%s N=%d M=%d" % (key, N, M)] = round(e0.elapsed_time(e1) / 10, 3)
    print(json.dumps(out))


if len(sys.argv) > 1:
    run()
else:
    res = {}
    for tag, env in (("cluster", {"MCPILCO_PERSIST": "1"}), ("fused", {"MCPILCO_NO_PERSIST": "1"})):
        o = subprocess.check_output([sys.executable, __file__, "x"], env=dict(os.environ, **env)).decode().strip().splitlines()[-1]
        res[tag] = json.loads(o)
    for k in res["cluster"]:
        print("%-18s cluster %.3f ms   fused %.3f ms   ratio %.2f" % (k, res["cluster"][k], res["fused"][k], res["cluster"][k] / res["fused"][k]))

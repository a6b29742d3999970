"""Debug aid: the whole-horizon cluster kernel against the per-step kernels on one small-N rollout, first step that differs."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch


def run(N):
    import native_helpers as nh
    from mcpilco_b200 import workloads as W
    sc = W.cartpole_sweep(N, nb=200, sigma_n=float(np.exp(-4.2)), se_only=False, seed=3)
    rs = np.random.RandomState(0)
    M, H = 64, 6
    sc.update(M=M, H=H)
    sc["eps0"] = rs.randn(M, sc["Ds"]); sc["eps"] = rs.randn(H - 1, M, sc["E"])
    sc["masks"] = (rs.rand(H, M, 200) >= sc["p_dropout"]).astype(np.float64)
    gps = nh.native_fit(sc)
    plan, _ = nh.native_plan(sc, gps, need_grad=True, inject=True, seed=1)
    plan.forward(nh.x0_of(sc)); torch.cuda.synchronize()
    return plan.states.cpu().numpy(), plan.inputs.cpu().numpy()


if len(sys.argv) > 2:
    s, u = run(int(sys.argv[1]))
    np.savez(sys.argv[2], s=s, u=u)
else:
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    for tag, env in (("persist", {}), ("perstep", {"MCPILCO_NO_SMALL_PATH": "1"})):
        subprocess.check_call([sys.executable, __file__, str(N), f"/tmp/pd_{tag}.npz"], env=dict(os.environ, **env))
    a, b = np.load("/tmp/pd_persist.npz"), np.load("/tmp/pd_perstep.npz")
    for t in range(a["s"].shape[0]):
        ds = np.abs(a["s"][t] - b["s"][t]); du = np.abs(a["u"][t] - b["u"][t])
        bad = np.argwhere(~(ds < 1e-9))
        print("t", t, "state err", np.nanmax(ds), "nan", int(np.isnan(a["s"][t]).sum()), "input err", np.nanmax(du), "nan", int(np.isnan(a["u"][t]).sum()),
              "first bad particles", sorted(set(bad[:, 0].tolist()))[:12])

"""Latency of one fwd+bwd rollout at the reference's REAL configuration shapes (C1-C4), CUDA path vs the oracle
port on the host CPU (1 thread = the reference's shipped setting, and all cores).  Synthetic fitted-like data of those shapes.
usage: python scripts/real_shapes.py [--no-cpu]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import native_helpers as nh
import scenarios
from mcpilco_b200 import workloads as W


def scale_scenario(name, N, M, H, nb, rs):
    """A tests/scenarios.py scenario re-drawn at real sizes."""
    if name in ("c1", "c2", "c3"):
        sc = W.cartpole_sweep(N, nb=nb, sigma_n=float(np.exp(-4.2)) if name != "c3" else 0.05, se_only=(name != "c1"), seed=3)
        if name == "c3":
            sc["model"]["T"] = 1.0 / 30
            sc["pms"] = {"std_pos": np.array([3e-3, 3e-3]), "pos_idx": [0, 2], "vel_idx": [1, 3], "fc": 0.5}
    else:  # UR5 shape: Ds=12, Du=6, E=6, D=24, SE + linear, trajectory policy/cost
        q = rs.uniform(-1.5, 1.5, (N, 6)); dq = rs.uniform(-2, 2, (N, 6)); u = rs.uniform(-1, 1, (N, 6))
        X = np.concatenate([dq, np.sin(q), np.cos(q), u], 1)
        Y = 0.02 * (3 * u - 2 * np.sin(q) - 0.3 * dq) + 0.005 * rs.randn(N, 6)
        gps = [{"log_ls": np.log(3.0) + 0.1 * rs.randn(24), "lambda": 1.0, "sigma_n": 0.05, "mean": 0.0, "mpk": [0.1 * np.exp(0.1 * rs.randn(25))]} for _ in range(6)]
        tt = np.linspace(0, 1, H)[:, None]
        traj = np.concatenate([0.3 * np.sin(2 * tt + np.arange(6)[None] * 0.3), 0.1 * np.cos(2 * tt + np.arange(6)[None] * 0.3)], 1)
        sc = dict(name="c4", D=24, Ds=12, Du=6, E=6, N=N, X=X, Y=Y, gps=gps,
                  model={"kind": "speed", "use_trig": True, "angle": list(range(6)), "not_angle": list(range(6, 12)), "vel": list(range(6, 12)),
                         "pos": list(range(6)), "T": 0.02},
                  policy={"kind": "target", "nb": nb, "centers": np.concatenate([np.pi / 2 * 2 * (rs.rand(nb, 12) - 0.5), 0.1 * 2 * (rs.rand(nb, 12) - 0.5)], 1),
                          "lengthscales": np.pi * np.ones(24), "weight": 2 * (rs.rand(6, nb) - 0.5), "u_max": [1.0] * 6, "target_traj": traj, "bias": None, "scale": None},
                  p_dropout=0.25, cost={"kind": "sat_traj", "target_traj": traj, "ls": np.array([0.5] * 6 + [1.0] * 6)},
                  x0_mean=traj[0].copy(), x0_var=1e-6 * np.ones(12))
    sc.update(M=M, H=H)
    sc["eps0"] = rs.randn(M, sc["Ds"]); sc["eps"] = rs.randn(H - 1, M, sc["E"])
    sc["masks"] = (rs.rand(H, M, nb) >= sc["p_dropout"]).astype(np.float64)
    if "pms" in sc:
        sc["meas_eps"] = rs.randn(H - 1, M, 2)
    return sc


LAST_FWD_MS = 0.0


def time_gpu(sc, reps=20):
    gps = nh.native_fit(sc)
    plan, _ = nh.native_plan(sc, gps, need_grad=True, inject=False, seed=1)
    x0 = nh.x0_of(sc)
    for _ in range(3):
        plan.forward(x0); plan.backward(grad_cost=1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        plan.forward(x0); plan.backward(grad_cost=1.0)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    total = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        plan.forward(x0)
    e1.record(); torch.cuda.synchronize()
    global LAST_FWD_MS
    LAST_FWD_MS = e0.elapsed_time(e1) / reps
    return total, wall * 1e3, float(plan.cost_out[0])


def time_cpu(sc, threads, reps=2):
    import helpers as Hh
    torch.set_num_threads(threads)
    gps = Hh.oracle_fit(sc)
    ts = []
    for _ in range(reps + 1):
        t0 = time.perf_counter(); Hh.oracle_rollout(sc, gps); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts[1:]))


def main():
    rs = np.random.RandomState(0)
    shapes = [("c1", 300, 400, 60, 200), ("c2", 300, 400, 60, 200), ("c3", 300, 400, 90, 200), ("c4", 400, 200, 200, 400), ("c1", 60, 400, 60, 200)]
    nsel = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--N=")]
    if nsel:
        shapes = [s for s in shapes if s[1] in nsel]
    only = [a.split("=")[1] for a in sys.argv if a.startswith("--only=")]
    reps = int(([a.split("=")[1] for a in sys.argv if a.startswith("--reps=")] or ["20"])[0])
    for name, N, M, H, nb in shapes:
        if only and name not in only:
            continue
        sc = scale_scenario(name, N, M, H, nb, rs)
        g_ms, g_wall, cost = time_gpu(sc, reps)
        row = {"config": name, "N": N, "M": M, "H": H, "nb": nb, "gpu_ms_fwd_bwd": round(g_ms, 3), "gpu_wall_ms": round(g_wall, 3),
               "gpu_ms_fwd": round(LAST_FWD_MS, 3), "gpu_ms_bwd": round(g_ms - LAST_FWD_MS, 3),
               "gpu_particle_steps_per_s": round(M * H / (g_ms * 1e-3)), "cost": cost}
        if "--no-cpu" not in sys.argv:
            c1 = time_cpu(sc, 1); call = time_cpu(sc, os.cpu_count())
            row.update(cpu_ms_1thread=round(c1, 1), cpu_ms_all=round(call, 1), cpu_threads=os.cpu_count(), speedup_vs_1thread=round(c1 / g_ms, 1))
        print(json.dumps(row), flush=True)

main()

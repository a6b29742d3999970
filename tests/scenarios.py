"""Seeded parity scenarios shared by the golden-vector generator (which drives the real reference),
the oracle tests and the GPU parity tests.  Pure numpy: no reference, oracle or CUDA import here.

Each scenario is a dict of numpy arrays / python scalars describing one small instance of the hot
path in the *reference's own parameterisation* (log-lengthscales, log MPK weights, ...), chosen
"fitted-like"  so that rollouts are smooth and gradients O(1-10).
"""
from __future__ import annotations

import numpy as np


def _cartpole_like_data(rs, N, sigma_n):
    p = rs.uniform(-2, 2, N); dp = rs.uniform(-5, 5, N); th = rs.uniform(-np.pi, np.pi, N)
    dth = rs.uniform(-10, 10, N); u = rs.uniform(-10, 10, N)
    X = np.stack([p, dp, dth, np.sin(th), np.cos(th), u], 1)
    # smooth synthetic velocity increments (not the true ODE: any smooth map will do for parity)
    y0 = 0.05 * (0.8 * u - 0.1 * dp + 0.5 * np.sin(th) * np.cos(th) + 0.02 * dth ** 2 * np.sin(th))
    y1 = 0.05 * (-3.0 * np.sin(th) - 1.2 * u * np.cos(th) + 0.05 * dp * np.cos(th))
    Y = np.stack([y0, y1], 1) + sigma_n * rs.randn(N, 2)
    return X, Y


def _policy_angles(rs, nb, u_max):
    ang = np.pi * 2 * (rs.rand(nb, 1) - 0.5)
    centers = np.concatenate([np.pi * 2 * (rs.rand(nb, 3) - 0.5), np.cos(ang), np.sin(ang)], 1)
    return {"kind": "angles", "nb": nb, "centers": centers, "lengthscales": np.ones(5) + 0.3 * rs.rand(5),
            "weight": u_max * (rs.rand(1, nb) - 0.5), "u_max": u_max, "angle": np.array([2]),
            "non_angle": np.array([0, 1, 3]), "bias": None, "scale": None}


def scenario(name, seed=0):
    rs = np.random.RandomState(1000 + seed)
    sc = {"name": name}
    if name in ("c1", "c2", "c3"):
        N, M, H, nb = 48, 12, 6, 20
        sigma_n = float(np.exp(-4.2)) if name != "c3" else 0.05
        X, Y = _cartpole_like_data(rs, N, sigma_n)
        gps = []
        for e in range(2):
            g = {"log_ls": np.array([2, 2, 2, 0.8, 1.5, 2.5]) + 0.1 * rs.randn(6), "lambda": 1.0,
                 "sigma_n": sigma_n, "mean": 0.0}
            if name == "c1":
                g["mpk"] = [np.exp(np.array([-5, -5, -5, -4, -4, -4, -3.0]) + 0.1 * rs.randn(7)),
                            np.exp(np.array([-5, -5, -4, -2, -1, -4.0] * 2) + 0.1 * rs.randn(12))]
            else:
                g["mpk"] = []
            gps.append(g)
        sc.update(D=6, Ds=4, Du=1, E=2, N=N, M=M, H=H, X=X, Y=Y, gps=gps,
                  model={"kind": "speed", "use_trig": True, "angle": [2], "not_angle": [0, 1, 3], "vel": [1, 3],
                         "pos": [0, 2], "T": 0.05 if name != "c3" else 1.0 / 30},
                  policy=_policy_angles(rs, nb, 10.0), p_dropout=0.25,
                  cost={"kind": "cart_pole", "target": np.array([np.pi, 0.0]), "ls": np.array([3.0, 1.0]),
                        "angle_index": 2, "pos_index": 0},
                  x0_mean=np.zeros(4), x0_var=1e-4 * np.ones(4))
        if name == "c3":
            sc["pms"] = {"std_pos": np.array([3e-3, 3e-3]), "pos_idx": [0, 2], "vel_idx": [1, 3], "fc": 0.5}
    elif name == "c4":  # UR5-like, two joints
        N, M, H, nb = 40, 10, 5, 16
        q = rs.uniform(-1.5, 1.5, (N, 2)); dq = rs.uniform(-2, 2, (N, 2)); u = rs.uniform(-1, 1, (N, 2))
        X = np.concatenate([dq, np.sin(q), np.cos(q), u], 1)
        Y = 0.02 * np.stack([3 * u[:, 0] - 2 * np.sin(q[:, 0]) - 0.3 * dq[:, 0],
                             3 * u[:, 1] - np.sin(q[:, 0] + q[:, 1]) - 0.3 * dq[:, 1]], 1) + 0.005 * rs.randn(N, 2)
        gps = [{"log_ls": np.log(3.0) + 0.1 * rs.randn(8), "lambda": 1.0, "sigma_n": 0.05, "mean": 0.0,
                "mpk": [0.1 * np.exp(0.1 * rs.randn(9))]} for _ in range(2)]
        tt = np.linspace(0, 1, H)[:, None]
        traj = np.concatenate([0.3 * np.sin(2 * tt + np.array([[0.0, 0.7]])), 0.1 * np.cos(2 * tt + np.array([[0.0, 0.7]]))], 1)
        sc.update(D=8, Ds=4, Du=2, E=2, N=N, M=M, H=H, X=X, Y=Y, gps=gps,
                  model={"kind": "speed", "use_trig": True, "angle": [0, 1], "not_angle": [2, 3], "vel": [2, 3],
                         "pos": [0, 1], "T": 0.02},
                  policy={"kind": "target", "nb": nb,
                          "centers": np.concatenate([np.pi / 2 * 2 * (rs.rand(nb, 4) - 0.5), 0.1 * 2 * (rs.rand(nb, 4) - 0.5)], 1),
                          "lengthscales": np.pi * np.ones(8), "weight": 2 * (rs.rand(2, nb) - 0.5),
                          "u_max": [1.0, 0.8], "target_traj": traj, "bias": None, "scale": None},
                  p_dropout=0.25,
                  cost={"kind": "sat_traj", "target_traj": traj, "ls": np.array([0.5, 0.5, 1.0, 1.0])},
                  x0_mean=traj[0].copy(), x0_var=1e-6 * np.ones(4))
    elif name == "delta":  # plain [x,u] features, delta-state model, plain policy with bias + scale, no squash
        N, M, H, nb = 36, 9, 5, 12
        X = rs.uniform(-1, 1, (N, 4))
        Y = 0.05 * np.stack([X[:, 1], -np.sin(2 * X[:, 0]) + X[:, 3], 0.5 * X[:, 0] * X[:, 2]], 1) + 0.01 * rs.randn(N, 3)
        gps = [{"log_ls": 0.5 + 0.1 * rs.randn(4), "lambda": 1.3, "sigma_n": 0.02, "mean": 0.01 * (e + 1), "mpk": []}
               for e in range(3)]
        sc.update(D=4, Ds=3, Du=1, E=3, N=N, M=M, H=H, X=X, Y=Y, gps=gps,
                  model={"kind": "delta", "use_trig": False, "angle": [], "not_angle": [0, 1, 2], "vel": [], "pos": [], "T": 0.0},
                  policy={"kind": "plain", "nb": nb, "centers": rs.uniform(-1, 1, (nb, 3)),
                          "lengthscales": 1.0 + rs.rand(3), "weight": rs.rand(1, nb) - 0.5, "u_max": None,
                          "bias": np.array([0.05]), "scale": np.array([1.0, 2.0, 0.5])},
                  p_dropout=0.1,
                  cost={"kind": "sat_target", "target": np.array([[0.2, -0.1]]), "ls": np.array([1.0, 2.0]), "active": [0, 2]},
                  x0_mean=np.array([0.1, 0.0, -0.1]), x0_var=1e-3 * np.ones(3))
    else:
        raise KeyError(name)
    M, H, Ds, E, nb = sc["M"], sc["H"], sc["Ds"], sc["E"], sc["policy"]["nb"]
    sc["eps0"] = rs.randn(M, Ds)
    sc["eps"] = rs.randn(H - 1, M, E)
    sc["masks"] = (rs.rand(H, M, nb) >= sc["p_dropout"]).astype(np.float64)
    if "pms" in sc:
        sc["meas_eps"] = rs.randn(H - 1, M, 2)
    return sc


def ur5_full(N=96, M=24, H=6, nb=40, seed=0):
    """Config 4 at its TRUE dimensions (Ds = 12, Du = 6, E = 6 outputs, D = 24 gp inputs, SE + linear kernel, trajectory policy and
    cost; test_mcpilco_ur5_mujoco.py:57-162) with small N / M / H.  No golden file: checked against the CPU oracle."""
    rs = np.random.RandomState(2000 + seed)
    q = rs.uniform(-1.5, 1.5, (N, 6)); dq = rs.uniform(-2, 2, (N, 6)); u = rs.uniform(-1, 1, (N, 6))
    X = np.concatenate([dq, np.sin(q), np.cos(q), u], 1)
    Y = 0.02 * (3 * u - 2 * np.sin(q) - 0.3 * dq) + 0.005 * rs.randn(N, 6)
    gps = [{"log_ls": np.log(3.0) + 0.1 * rs.randn(24), "lambda": 1.0, "sigma_n": 0.05, "mean": 0.0, "mpk": [0.1 * np.exp(0.1 * rs.randn(25))]}
           for _ in range(6)]
    tt = np.linspace(0, 1, H)[:, None]
    traj = np.concatenate([0.3 * np.sin(2 * tt + np.arange(6)[None] * 0.3), 0.1 * np.cos(2 * tt + np.arange(6)[None] * 0.3)], 1)
    sc = dict(name="ur5", D=24, Ds=12, Du=6, E=6, N=N, M=M, H=H, X=X, Y=Y, gps=gps,
              model={"kind": "speed", "use_trig": True, "angle": list(range(6)), "not_angle": list(range(6, 12)), "vel": list(range(6, 12)),
                     "pos": list(range(6)), "T": 0.02},
              policy={"kind": "target", "nb": nb, "centers": np.concatenate([np.pi / 2 * 2 * (rs.rand(nb, 12) - 0.5), 0.1 * 2 * (rs.rand(nb, 12) - 0.5)], 1),
                      "lengthscales": np.pi * np.ones(24), "weight": 2 * (rs.rand(6, nb) - 0.5), "u_max": [1.0] * 6, "target_traj": traj,
                      "bias": None, "scale": None},
              p_dropout=0.25, cost={"kind": "sat_traj", "target_traj": traj, "ls": np.array([0.5] * 6 + [1.0] * 6)},
              x0_mean=traj[0].copy(), x0_var=1e-6 * np.ones(12))
    sc["eps0"] = rs.randn(M, 12); sc["eps"] = rs.randn(H - 1, M, 6)
    sc["masks"] = (rs.rand(H, M, nb) >= 0.25).astype(np.float64)
    return sc


def headline(N=8192, M=256, H=3, nb=50, seed=0):
    """The bench workload's shape family (BASELINE.json configs[1]: cart-pole, SE + MPK(2) kernels, E = 2, D = 6) at its FULL training
    size with few particles and a short horizon, so that the CPU oracle finishes in seconds.  No golden file: checked against the oracle."""
    rs = np.random.RandomState(3000 + seed)
    sc = scenario("c1", seed)
    X, Y = _cartpole_like_data(rs, N, 0.1)
    for g in sc["gps"]:
        g["sigma_n"] = 0.1
    sc.update(name="headline", N=N, M=M, H=H, X=X, Y=Y, policy=_policy_angles(rs, nb, 10.0))
    sc["eps0"] = rs.randn(M, 4); sc["eps"] = rs.randn(H - 1, M, 2)
    sc["masks"] = (rs.rand(H, M, nb) >= sc["p_dropout"]).astype(np.float64)
    return sc


ALL = ("c1", "c2", "c3", "c4", "delta")

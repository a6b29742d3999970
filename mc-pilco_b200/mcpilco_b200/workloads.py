"""Synthetic workloads of the named benchmark shapes.  Pure numpy, seeded; no CUDA, no oracle import.
bench.py hands the SAME arrays to the CUDA arm and to the CPU baseline arm.

C5 ("synthetic cartpole GP scaling sweep"): cart-pole transitions — states p~U(-2,2), dp~U(-5,5), theta~U(-pi,pi),
dtheta~U(-10,10), u~U(-10,10); targets = velocity increments after one RK4 step (dt = 0.05 s) of the cart-pole ODE
(constants of the reference's simulation_class/ode_systems.py:43-66) plus N(0, sigma_n^2); gp-input rows
[p, dp, dtheta, sin theta, cos theta, u].  Hyper-parameters fixed, fitted-like: SE log-lengthscales [2,2,2,0.8,1.5,2.5],
lambda 1, MPK_1 log s = [-5,-5,-5,-4,-4,-4,-3], MPK_2 log p = [-5,-5,-4,-2,-1,-4] x 2, sigma_n = 0.1.
Policy / cost / initial distribution as in test_mcpilco_cartpole.py:137-146: nb = 200 squashed RBF policy (u_max 10,
p_dropout 0.25), Cart_pole_cost(target [pi, 0], lengthscales [3, 1]), x0 ~ N(0, 1e-4 I).
"""
import numpy as np

SE_LOG_LS = np.array([2.0, 2.0, 2.0, 0.8, 1.5, 2.5])
MPK1_LOG = np.array([-5.0, -5.0, -5.0, -4.0, -4.0, -4.0, -3.0])
MPK2_LOG = np.array([-5.0, -5.0, -4.0, -2.0, -1.0, -4.0] * 2)


def _cartpole_acc(s, u):
    mc, mp, ll, g, bb = 0.5, 0.5, 0.5, 9.81, 0.1
    dp, th, dth = s[:, 1], s[:, 2], s[:, 3]
    st, ct = np.sin(th), np.cos(th)
    den = 4 * (mc + mp) - 3 * mp * ct ** 2
    ddp = (2 * mp * ll * dth ** 2 * st + 3 * mp * g * st * ct + 4 * u - 4 * bb * dp) / den
    ddth = (-3 * mp * ll * dth ** 2 * st * ct - 6 * (mc + mp) * g * st - 6 * (u - bb * dp) * ct) / (ll * den)
    return np.stack([dp, ddp, dth, ddth], 1)


def cartpole_transitions(N, sigma_n, rs):
    r = rs.rand(N, 5)
    s = np.stack([4 * r[:, 0] - 2, 10 * r[:, 1] - 5, 2 * np.pi * r[:, 2] - np.pi, 20 * r[:, 3] - 10], 1)
    u = 20 * r[:, 4] - 10
    dt = 0.05
    k1 = _cartpole_acc(s, u); k2 = _cartpole_acc(s + dt / 2 * k1, u)
    k3 = _cartpole_acc(s + dt / 2 * k2, u); k4 = _cartpole_acc(s + dt * k3, u)
    s1 = s + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    X = np.stack([s[:, 0], s[:, 1], s[:, 3], np.sin(s[:, 2]), np.cos(s[:, 2]), u], 1)
    Y = np.stack([s1[:, 1] - s[:, 1], s1[:, 3] - s[:, 3]], 1) + sigma_n * rs.randn(N, 2)
    return X, Y


def cartpole_sweep(N, nb=200, sigma_n=0.1, se_only=False, seed=0):
    """Scenario dict in the format of tests/scenarios.py (without M/H-dependent noise)."""
    rs = np.random.RandomState(seed)
    X, Y = cartpole_transitions(N, sigma_n, rs)
    gps = [{"log_ls": SE_LOG_LS.copy(), "lambda": 1.0, "sigma_n": sigma_n, "mean": 0.0,
            "mpk": [] if se_only else [np.exp(MPK1_LOG), np.exp(MPK2_LOG)]} for _ in range(2)]
    ang = np.pi * 2 * (rs.rand(nb, 1) - 0.5)
    policy = {"kind": "angles", "nb": nb, "centers": np.concatenate([np.pi * 2 * (rs.rand(nb, 3) - 0.5), np.cos(ang), np.sin(ang)], 1),
              "lengthscales": np.ones(5), "weight": 10.0 * (rs.rand(1, nb) - 0.5), "u_max": 10.0, "angle": np.array([2]),
              "non_angle": np.array([0, 1, 3]), "bias": None, "scale": None}
    return {"name": "c5", "D": 6, "Ds": 4, "Du": 1, "E": 2, "N": N, "X": X, "Y": Y, "gps": gps,
            "model": {"kind": "speed", "use_trig": True, "angle": [2], "not_angle": [0, 1, 3], "vel": [1, 3], "pos": [0, 2], "T": 0.05},
            "policy": policy, "p_dropout": 0.25,
            "cost": {"kind": "cart_pole", "target": np.array([np.pi, 0.0]), "ls": np.array([3.0, 1.0]), "angle_index": 2, "pos_index": 0},
            "x0_mean": np.zeros(4), "x0_var": 1e-4 * np.ones(4)}


def flops_per_particle_step(N_list, D, need_grad=True):
    """ALGORITHMIC flops of one particle-step: sum_i 2 N_i^2 + (20 D + 20) N_i with the backward pass,
    sum_i N_i^2 + (10 D + 10) N_i forward only."""
    if need_grad:
        return float(sum(2.0 * n * n + (20.0 * D + 20.0) * n for n in N_list))
    return float(sum(1.0 * n * n + (10.0 * D + 10.0) * n for n in N_list))

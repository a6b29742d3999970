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


REAL_SHAPES = {  # name -> (config, N, M, H, nb): the reference's own sizes (C1-C4)
    "c1": ("c1", 300, 400, 60, 200),        # test_mcpilco_cartpole.py:124,199 (SE + MPK(2), SoD-sized training set of a late trial)
    "c2": ("c2", 300, 400, 60, 200),        # test_mcpilco_cartpole_rbf_ker.py (SE only)
    "c3": ("c3", 300, 400, 90, 200),        # test_mcpilco4pms_cartpole.py:104,171 (4PMS, T = 1/30 s -> H = 90)
    "c4": ("c4", 400, 200, 200, 400),       # test_mcpilco_ur5_mujoco.py:127,195 (D = 24, E = 6, Ds = 12, Du = 6)
    "c1_first_trial": ("c1", 60, 400, 60, 200),
}


def ur5_like(N, H, nb, rs):
    """UR5 joint-space shape (config 4): Ds = 12, Du = 6, E = 6 outputs, D = 24 gp inputs [dq, sin q, cos q, u], SE + linear kernel,
    trajectory-tracking policy and cost, on smooth synthetic second-order dynamics."""
    q = rs.uniform(-1.5, 1.5, (N, 6)); dq = rs.uniform(-2, 2, (N, 6)); u = rs.uniform(-1, 1, (N, 6))
    X = np.concatenate([dq, np.sin(q), np.cos(q), u], 1)
    Y = 0.02 * (3 * u - 2 * np.sin(q) - 0.3 * dq) + 0.005 * rs.randn(N, 6)
    gps = [{"log_ls": np.log(3.0) + 0.1 * rs.randn(24), "lambda": 1.0, "sigma_n": 0.05, "mean": 0.0, "mpk": [0.1 * np.exp(0.1 * rs.randn(25))]}
           for _ in range(6)]
    tt = np.linspace(0, 1, H)[:, None]
    traj = np.concatenate([0.3 * np.sin(2 * tt + np.arange(6)[None] * 0.3), 0.1 * np.cos(2 * tt + np.arange(6)[None] * 0.3)], 1)
    return dict(name="c4", D=24, Ds=12, Du=6, E=6, N=N, X=X, Y=Y, gps=gps,
                model={"kind": "speed", "use_trig": True, "angle": list(range(6)), "not_angle": list(range(6, 12)), "vel": list(range(6, 12)),
                       "pos": list(range(6)), "T": 0.02},
                policy={"kind": "target", "nb": nb,
                        "centers": np.concatenate([np.pi / 2 * 2 * (rs.rand(nb, 12) - 0.5), 0.1 * 2 * (rs.rand(nb, 12) - 0.5)], 1),
                        "lengthscales": np.pi * np.ones(24), "weight": 2 * (rs.rand(6, nb) - 0.5), "u_max": [1.0] * 6, "target_traj": traj,
                        "bias": None, "scale": None},
                p_dropout=0.25, cost={"kind": "sat_traj", "target_traj": traj, "ls": np.array([0.5] * 6 + [1.0] * 6)},
                x0_mean=traj[0].copy(), x0_var=1e-6 * np.ones(12))


def real_shape(key, seed=0, N=None, M=None, H=None, nb=None, with_noise=True):
    """A scenario dict (tests/scenarios.py format) at one of the reference's REAL configuration sizes (REAL_SHAPES), synthetic
    fitted-like data; `with_noise` adds the injected-noise tensors (eps0, eps, masks, meas_eps) the parity tests feed to both sides."""
    name, N0, M0, H0, nb0 = REAL_SHAPES[key]
    N, M, H, nb = N or N0, M or M0, H or H0, nb or nb0
    rs = np.random.RandomState(7000 + seed)
    if name in ("c1", "c2", "c3"):
        sc = cartpole_sweep(N, nb=nb, sigma_n=float(np.exp(-4.2)) if name != "c3" else 0.05, se_only=(name != "c1"), seed=3 + seed)
        sc["name"] = name
        if name == "c3":
            sc["model"]["T"] = 1.0 / 30
            sc["pms"] = {"std_pos": np.array([3e-3, 3e-3]), "pos_idx": [0, 2], "vel_idx": [1, 3], "fc": 0.5}
    else:
        sc = ur5_like(N, H, nb, rs)
    sc.update(M=M, H=H)
    if with_noise:
        sc["eps0"] = rs.randn(M, sc["Ds"]); sc["eps"] = rs.randn(H - 1, M, sc["E"])
        sc["masks"] = (rs.rand(H, M, nb) >= sc["p_dropout"]).astype(np.float64)
        if "pms" in sc:
            sc["meas_eps"] = rs.randn(H - 1, M, 2)
    return sc


def flops_per_particle_step(N_list, D, need_grad=True):
    """ALGORITHMIC flops of one particle-step: sum_i 2 N_i^2 + (20 D + 20) N_i with the backward pass,
    sum_i N_i^2 + (10 D + 10) N_i forward only."""
    if need_grad:
        return float(sum(2.0 * n * n + (20.0 * D + 20.0) * n for n in N_list))
    return float(sum(1.0 * n * n + (10.0 * D + 10.0) * n for n in N_list))


def build_pilco(sc, dev, pretrain=True):
    """The reference's construction sequence (test_mcpilco_cartpole.py:49-231, test_mcpilco4pms_cartpole.py, test_mcpilco_ur5_mujoco.py:57-162)
    against mcpilco_b200's classes for any scenario dict of this module / tests/scenarios.py: model-learning object with the scenario's
    GP hyper-parameters and training set, policy, cost, MC_PILCO / MC_PILCO4PMS object.  Returns the MC_PILCO object."""
    import contextlib
    import sys

    import torch

    from .model_learning import Model_learning as ML
    from .policy_learning import Cost_function as CF
    from .policy_learning import MC_PILCO as MCP
    from .policy_learning import Policy as PO
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)  # noqa: E731
    D = sc["D"]
    dicts = []
    for g in sc["gps"]:
        rbf = dict(active_dims=np.arange(D), lengthscales_init=np.exp(g["log_ls"]), flg_train_lengthscales=True, lambda_init=np.array([g["lambda"]]),
                   flg_train_lambda=False, sigma_n_init=np.array([g["sigma_n"]]), flg_train_sigma_n=True, mean_init=np.array([g["mean"]]),
                   sigma_n_num=None, dtype=torch.float64, device=dev)
        if g["mpk"]:
            mpk = dict(active_dims=np.arange(D), poly_deg=len(g["mpk"]), Sigma_pos_par_init_list=list(g["mpk"]),
                       flg_train_Sigma_pos_par_list=[True] * len(g["mpk"]), dtype=torch.float64, device=dev)
            dicts.append([rbf, mpk])
        else:
            dicts.append(rbf)
    m, p, c = sc["model"], sc["policy"], sc["cost"]
    if m["kind"] == "speed":
        cls = ML.Speed_Model_learning_RBF_MPK_angle_state if sc["gps"][0]["mpk"] else ML.Speed_Model_learning_RBF_angle_state
        ml = cls(num_gp=sc["E"], init_dict_list=dicts, T_sampling=m["T"], angle_indeces=m["angle"], not_angle_indeces=m["not_angle"],
                 vel_indeces=m["vel"], not_vel_indeces=m["pos"], device=dev)
    else:
        ml = ML.Model_learning_RBF(num_gp=sc["E"], init_dict_list=dicts, device=dev)
    ml.gp_inputs = T(sc["X"])
    ml.gp_output_list = [T(sc["Y"][:, e:e + 1]) for e in range(sc["E"])]
    ml.dim_state, ml.dim_input, ml.num_samples = sc["Ds"], sc["Du"], sc["N"]
    if pretrain:
        with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
            for e in range(sc["E"]):
                ml.pretrain_gp(e)
        ml.set_eval_mode()
    pkw = dict(input_dim=sc["Du"], num_basis=p["nb"], lengthscales_init=p["lengthscales"], centers_init=p["centers"], weight_init=p["weight"],
               flg_squash=p["u_max"] is not None, u_max=p["u_max"] if p["u_max"] is not None else 1, flg_drop=True, flg_bias=p["bias"] is not None,
               bias_init=p["bias"], device=dev)
    if p["kind"] == "angles":
        pcls, pkw = PO.Sum_of_gaussians_with_angles, dict(pkw, state_dim=sc["Ds"], angle_indices=p["angle"], non_angle_indices=p["non_angle"])
    elif p["kind"] == "target":
        pcls, pkw = PO.Sum_of_gaussians_with_target_trajectory, dict(pkw, state_dim=2 * sc["Ds"], target_traj=p["target_traj"])
    else:
        pcls, pkw = PO.Sum_of_gaussians, dict(pkw, state_dim=sc["Ds"], scale_factor=p["scale"])
    if c["kind"] == "cart_pole":
        ccls, ckw = CF.Cart_pole_cost, dict(target_state=T(c["target"]), lengthscales=T(c["ls"]), angle_index=c["angle_index"], pos_index=c["pos_index"])
    elif c["kind"] == "sat_traj":
        ccls, ckw = CF.Expected_saturated_distance_from_trajectory, dict(target_traj=T(c["target_traj"]), lengthscales=T(c["ls"]))
    else:
        ccls, ckw = CF.Expected_saturated_distance, dict(target_state=T(c["target"]), lengthscales=T(c["ls"]), active_dims=c["active"])
    common = dict(T_sampling=m["T"] or 0.05, state_dim=sc["Ds"], input_dim=sc["Du"], f_sim=None, f_model_learning=lambda: ml, model_learning_par={},
                  f_rand_exploration_policy=None, rand_exploration_policy_par=None, f_control_policy=pcls, control_policy_par=pkw,
                  f_cost_function=ccls, cost_function_par=ckw, device=dev)
    if "pms" in sc:
        q = sc["pms"]
        std = np.zeros(sc["Ds"])
        std[list(q["pos_idx"])] = q["std_pos"]
        return MCP.MC_PILCO4PMS(pos_indeces=q["pos_idx"], vel_indeces=q["vel_idx"], filtering_dict={"fc": q["fc"]}, std_meas_noise=std, **common)
    return MCP.MC_PILCO(**common)


def apply_kwargs(sc, dev, num_particles=None):
    import torch
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)  # noqa: E731
    return dict(particles_initial_state_mean=T(sc["x0_mean"]), particles_initial_state_var=T(sc["x0_var"]), flg_particles_init_uniform=False,
                particles_init_up_bound=None, particles_init_low_bound=None, flg_particles_init_multi_gauss=False,
                num_particles=int(num_particles or sc["M"]), T_control=sc["H"], p_dropout=sc["p_dropout"])

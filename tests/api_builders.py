"""Scenario -> objects of the reference's CLASS API.  The same construction code is run against the reference's own modules
(tests/golden/make_golden.py, in the build container) and against mcpilco_b200's mirror of them (tests/test_gpu_api.py):
`R` is a namespace with ML / PO / CF / MCP modules.  That the code is shared is the drop-in claim being tested."""
import numpy as np
import torch


def tensor_factory(device):
    return lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device=device)


def build_model(R, sc, device, pretrain=True, approximation=None):
    T = tensor_factory(device)
    D = sc["D"]
    dicts = []
    for g in sc["gps"]:
        rbf = dict(active_dims=np.arange(D), lengthscales_init=np.exp(g["log_ls"]), flg_train_lengthscales=True,
                   lambda_init=np.array([g["lambda"]]), flg_train_lambda=False, sigma_n_init=np.array([g["sigma_n"]]),
                   flg_train_sigma_n=True, mean_init=np.array([g["mean"]]), sigma_n_num=None, dtype=torch.float64, device=device)
        if g["mpk"]:
            mpk = dict(active_dims=np.arange(D), poly_deg=len(g["mpk"]), Sigma_pos_par_init_list=list(g["mpk"]),
                       flg_train_Sigma_pos_par_list=[True] * len(g["mpk"]), dtype=torch.float64, device=device)
            dicts.append([rbf, mpk])
        else:
            dicts.append(rbf)
    m = sc["model"]
    has_mpk = bool(sc["gps"][0]["mpk"])
    extra = {} if approximation is None else dict(approximation_mode="SOD", approximation_dict=approximation)
    if m["kind"] == "speed":
        cls = R.ML.Speed_Model_learning_RBF_MPK_angle_state if has_mpk else R.ML.Speed_Model_learning_RBF_angle_state
        ml = cls(num_gp=sc["E"], init_dict_list=dicts, T_sampling=m["T"], angle_indeces=m["angle"], not_angle_indeces=m["not_angle"],
                 vel_indeces=m["vel"], not_vel_indeces=m["pos"], device=device, **extra)
    else:
        ml = R.ML.Model_learning_RBF(num_gp=sc["E"], init_dict_list=dicts, device=device, **extra)
    ml.gp_inputs = T(sc["X"])
    ml.gp_output_list = [T(sc["Y"][:, e:e + 1]) for e in range(sc["E"])]
    ml.dim_state, ml.dim_input, ml.num_samples = sc["Ds"], sc["Du"], sc["N"]
    if pretrain:
        with torch.no_grad():
            for e in range(sc["E"]):
                ml.pretrain_gp(e)
        ml.set_eval_mode()
    return ml


def policy_kwargs(sc, device):
    p = sc["policy"]
    kw = dict(input_dim=sc["Du"], num_basis=p["nb"], lengthscales_init=p["lengthscales"], centers_init=p["centers"],
              weight_init=p["weight"], flg_squash=p["u_max"] is not None, u_max=p["u_max"] if p["u_max"] is not None else 1,
              flg_drop=True, flg_bias=p["bias"] is not None, bias_init=p["bias"], device=device)
    if p["kind"] == "angles":
        kw.update(state_dim=sc["Ds"], angle_indices=p["angle"], non_angle_indices=p["non_angle"])
    elif p["kind"] == "target":
        kw.update(state_dim=2 * sc["Ds"], target_traj=p["target_traj"])
    else:
        kw.update(state_dim=sc["Ds"], scale_factor=p["scale"])
    return kw


def build_policy(R, sc, device):
    T = tensor_factory(device)
    cls = {"angles": R.PO.Sum_of_gaussians_with_angles, "target": R.PO.Sum_of_gaussians_with_target_trajectory,
           "plain": R.PO.Sum_of_gaussians}[sc["policy"]["kind"]]
    pol = cls(**policy_kwargs(sc, device))
    if sc["policy"]["bias"] is not None:
        pol.f_linear.bias.data = T(sc["policy"]["bias"])
    return pol


def cost_spec(R, sc, device):
    T = tensor_factory(device)
    c = sc["cost"]
    if c["kind"] == "cart_pole":
        return R.CF.Cart_pole_cost, dict(target_state=T(c["target"]), lengthscales=T(c["ls"]), angle_index=c["angle_index"], pos_index=c["pos_index"])
    if c["kind"] == "sat_traj":
        return R.CF.Expected_saturated_distance_from_trajectory, dict(target_traj=T(c["target_traj"]), lengthscales=T(c["ls"]))
    if c["kind"] == "sat_target":
        return R.CF.Expected_saturated_distance, dict(target_state=T(c["target"]), lengthscales=T(c["ls"]), active_dims=c["active"])
    raise KeyError(c["kind"])


def build_pilco(R, sc, ml, device, rand_policy=None):
    cost_cls, cost_par = cost_spec(R, sc, device)
    common = dict(T_sampling=sc["model"]["T"] or 0.05, state_dim=sc["Ds"], input_dim=sc["Du"], f_sim=None,
                  f_model_learning=lambda: ml, model_learning_par={}, f_rand_exploration_policy=rand_policy,
                  rand_exploration_policy_par=dict(state_dim=sc["Ds"], input_dim=sc["Du"]),
                  f_control_policy=lambda: build_policy(R, sc, device), control_policy_par={}, f_cost_function=cost_cls,
                  cost_function_par=cost_par, device=device)
    if "pms" in sc:
        return R.MCP.MC_PILCO4PMS(pos_indeces=sc["pms"]["pos_idx"], vel_indeces=sc["pms"]["vel_idx"], filtering_dict={"fc": sc["pms"]["fc"]},
                                  std_meas_noise=np.array([sc["pms"]["std_pos"][0], 0.0, sc["pms"]["std_pos"][1], 0.0]), **common)
    return R.MCP.MC_PILCO(**common)


def apply_kwargs(sc, device):
    T = tensor_factory(device)
    return dict(particles_initial_state_mean=T(sc["x0_mean"]), particles_initial_state_var=T(sc["x0_var"]), flg_particles_init_uniform=False,
                particles_init_up_bound=None, particles_init_low_bound=None, flg_particles_init_multi_gauss=False, num_particles=sc["M"],
                T_control=sc["H"], p_dropout=sc["p_dropout"])

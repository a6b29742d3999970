#!/usr/bin/env python
"""A miniature MC-PILCO trial loop on the GPU, written against mcpilco_b200's mirror of the reference's classes — the same sequence
`MC_PILCO.reinforce` drives (reference policy_learning/MC_PILCO.py:89-258): collect data from the system, train the GP model
(`reinforce_model`), optimise the policy on particle rollouts (`reinforce_policy`), apply it, repeat.

The "system" here is a small RK4 cart-pole simulator in numpy (the reference uses scipy odeint / MuJoCo; simulators are outside this
repository's scope), the exploration input is a sum of sinusoids.  Everything between data collection and the learned policy —
kernel matrices, Cholesky, hyper-parameter gradients, particle rollouts, backprop through time — runs in libmcpilco_b200.so.

    python examples/cartpole_swingup.py [--trials 2] [--opt-steps 150] [--particles 400]
"""
import argparse
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200"))

import numpy as np
import torch

import mcpilco_b200.gpr_lib.Likelihood.Gaussian_likelihood as Likelihood
import mcpilco_b200.model_learning.Model_learning as ML
import mcpilco_b200.policy_learning.Cost_function as Cost_function
import mcpilco_b200.policy_learning.MC_PILCO as MC_PILCO
import mcpilco_b200.policy_learning.Policy as Policy
from mcpilco_b200.workloads import _cartpole_acc

T_SAMPLING, T_CONTROL, U_MAX = 0.05, 3.0, 10.0


def simulate(policy_fn, x0, steps, rs, noise=1e-2):
    """RK4 cart-pole, zero-order-hold input; returns noisy states [steps, 4] and inputs [steps, 1] (state = [p, dp, theta, dtheta])."""
    xs, us, x = [], [], np.array(x0, dtype=np.float64)
    for t in range(steps):
        meas = x + noise * rs.randn(4)
        u = float(np.clip(policy_fn(meas, t), -U_MAX, U_MAX))
        xs.append(meas); us.append([u])
        f = lambda s: _cartpole_acc(s[None, :], np.array([u]))[0]  # noqa: E731
        h = T_SAMPLING / 4
        for _ in range(4):
            k1 = f(x); k2 = f(x + h / 2 * k1); k3 = f(x + h / 2 * k2); k4 = f(x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return np.array(xs), np.array(us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=2)
    ap.add_argument("--opt-steps", type=int, default=150)
    ap.add_argument("--particles", type=int, default=400)
    ap.add_argument("--gp-epochs", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the hot path has no CPU fallback")
    dev, dtype = torch.device("cuda:0"), torch.float64
    rs = np.random.RandomState(args.seed)
    torch.manual_seed(args.seed)

    # ---- the reference's configuration dicts (test_mcpilco_cartpole_rbf_ker.py: squared-exponential GPs), device = cuda ----
    rbf = dict(active_dims=np.arange(6), lengthscales_init=np.ones(6), flg_train_lengthscales=True, lambda_init=np.ones(1), flg_train_lambda=False,
               sigma_n_init=np.ones(1), flg_train_sigma_n=True, dtype=dtype, device=dev)
    model_par = dict(num_gp=2, init_dict_list=[rbf] * 2, T_sampling=T_SAMPLING, angle_indeces=[2], not_angle_indeces=[0, 1, 3],
                     vel_indeces=[1, 3], not_vel_indeces=[0, 2], device=dev, dtype=dtype,
                     approximation_mode="SOD", approximation_dict={"SOD_threshold_mode": "relative", "SOD_threshold": 0.5, "flg_SOD_permutation": False})
    nb = 200
    ang = np.pi * 2 * (rs.rand(nb, 1) - 0.5)
    policy_par = dict(state_dim=4, input_dim=1, num_basis=nb, angle_indices=np.array([2]), non_angle_indices=np.array([0, 1, 3]),
                      lengthscales_init=np.ones(5), centers_init=np.concatenate([np.pi * 2 * (rs.rand(nb, 3) - 0.5), np.cos(ang), np.sin(ang)], 1),
                      weight_init=U_MAX * (rs.rand(1, nb) - 0.5), flg_squash=True, u_max=U_MAX, flg_drop=True, dtype=dtype, device=dev)
    cost_par = dict(target_state=torch.tensor([np.pi, 0.0], dtype=dtype, device=dev), lengthscales=torch.tensor([3.0, 1.0], dtype=dtype, device=dev),
                    angle_index=2, pos_index=0)
    obj = MC_PILCO.MC_PILCO(T_sampling=T_SAMPLING, state_dim=4, input_dim=1, f_sim=None, f_model_learning=ML.Speed_Model_learning_RBF_angle_state,
                            model_learning_par=model_par, f_rand_exploration_policy=None, rand_exploration_policy_par=None,
                            f_control_policy=Policy.Sum_of_gaussians_with_angles, control_policy_par=policy_par,
                            f_cost_function=Cost_function.Cart_pole_cost, cost_function_par=cost_par, dtype=dtype, device=dev)
    gp_opt = {"f_optimizer": "lambda p : torch.optim.Adam(p, lr=0.01)", "criterion": Likelihood.Marginal_log_likelihood, "N_epoch": args.gp_epochs,
              "N_epoch_print": 10 ** 9}
    steps = int(T_CONTROL / T_SAMPLING)
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731

    # ---- exploration trajectory: sum of sinusoids ----
    amp, freq, ph = U_MAX * rs.rand(5) / 2, 2 * np.pi * rs.rand(5) * 1.5, 2 * np.pi * rs.rand(5)
    xs, us = simulate(lambda x, t: float(np.sum(amp * np.sin(freq * t * T_SAMPLING + ph))), np.zeros(4), steps, rs)
    obj.model_learning.add_data(xs, us)
    for trial in range(args.trials):
        t0 = time.time()
        with quiet():
            obj.model_learning.reinforce_model(optimization_opt_list=[gp_opt] * 2)
        obj.model_learning.set_eval_mode()
        t_model = time.time() - t0
        sizes = [int(x.shape[0]) for x in obj.model_learning.gp_inputs_tr_list]
        t0 = time.time()
        with quiet():
            cost_list, std_list, st, inp = obj.reinforce_policy(
                T_control=T_CONTROL, num_particles=args.particles, trial_index=trial,
                particles_initial_state_mean=torch.zeros(4, dtype=dtype, device=dev), particles_initial_state_var=1e-4 * torch.ones(4, dtype=dtype, device=dev),
                flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None, flg_particles_init_multi_gauss=False,
                opt_steps_list=[args.opt_steps] * args.trials, lr_list=[0.01] * args.trials, f_optimizer="lambda p, lr : torch.optim.Adam(p, lr)",
                num_step_print=10 ** 9, p_dropout_list=[0.25] * args.trials, p_drop_reduction=0.125, min_step=200, num_min_diff_cost=200,
                policy_reinit_dict=dict(lenghtscales_par=np.ones(5), centers_par=np.array([np.pi, np.pi, np.pi, 1.0, 1.0]), weight_par=U_MAX), max_reinit=3)
        torch.cuda.synchronize()
        t_policy = time.time() - t0
        pol = obj.control_policy
        with torch.no_grad():
            xs, us = simulate(lambda x, t: float(pol(torch.tensor(x[None, :], dtype=dtype, device=dev), t=t, p_dropout=0.0)[0, 0]), np.zeros(4), steps, rs)
        obj.model_learning.set_training_mode()
        obj.model_learning.add_data(xs, us)
        print("trial %d: GP training sets %s, model update %.1f s | policy: %d steps in %.1f s (%.1f ms/step), particle cost %.2f -> %.2f | "
              "system: final |theta| %.2f rad, final p %.2f m" % (trial, sizes, t_model, len(cost_list), t_policy, 1e3 * t_policy / max(len(cost_list), 1),
                                                                  cost_list[0], cost_list[-1], abs(xs[-1, 2]), xs[-1, 0]), flush=True)


if __name__ == "__main__":
    main()

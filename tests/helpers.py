"""Scenario -> oracle objects (torch CPU fp64).  Test-only."""
import math
import os

import numpy as np
import torch

from oracle import mcpilco_oracle as O

T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def oracle_specs(sc):
    return [O.make_spec(sc["D"], log_ls=g["log_ls"], log_lambda=math.log(g["lambda"]), mean=g["mean"],
                        mpk_log_pars=[np.log(w) for w in g["mpk"]], sigma_n=g["sigma_n"]) for g in sc["gps"]]


def oracle_model(sc):
    m = dict(sc["model"])
    m.update(Ds=sc["Ds"], Du=sc["Du"], norm=[1.0] * sc["E"])
    return m


def oracle_policy(sc, requires_grad=False):
    p = sc["policy"]
    Dp = p["centers"].shape[1]
    pol = {"kind": p["kind"], "log_ls": T(np.log(p["lengthscales"])).reshape(1, -1), "centers": T(p["centers"]),
           "W": T(p["weight"]), "bias": None if p["bias"] is None else T(p["bias"]), "u_max": p["u_max"],
           "scale": T(np.ones(Dp) if p.get("scale") is None else p["scale"]).reshape(1, -1)}
    if p["kind"] == "angles":
        pol.update(angle=list(p["angle"]), non_angle=list(p["non_angle"]))
    if p["kind"] == "target":
        pol["target_traj"] = T(p["target_traj"])
    if requires_grad:
        for k in ("log_ls", "centers", "W", "bias"):
            if pol[k] is not None:
                pol[k].requires_grad_(True)
    return pol


def oracle_fit(sc):
    specs = oracle_specs(sc)
    X = T(sc["X"])
    gps = []
    for e, sp in enumerate(specs):
        alpha, _, Kinv = O.gp_fit(sp, X, T(sc["Y"][:, e:e + 1]))
        gps.append((sp, X, alpha, Kinv))
    return gps


def oracle_cost(sc, states):
    c = sc["cost"]
    if c["kind"] == "cart_pole":
        return O.cost_cart_pole(states, T(c["target"]), T(c["ls"]), c["angle_index"], c["pos_index"])
    if c["kind"] == "sat_traj":
        return O.cost_saturated_trajectory(states, T(c["target_traj"]), T(c["ls"]))
    if c["kind"] == "sat_target":
        return O.cost_saturated_distance(states, T(c["target"]), T(c["ls"]), c["active"])
    raise KeyError(c["kind"])


def oracle_rollout(sc, gps=None, requires_grad=True):
    """Full oracle rollout + cost (+ grads).  Returns dict like the golden file."""
    gps = gps or oracle_fit(sc)
    model = oracle_model(sc)
    pol = oracle_policy(sc, requires_grad)
    x0 = O.initial_particles(T(sc["x0_mean"]), T(sc["x0_var"]), T(sc["eps0"]))
    if "pms" in sc:
        q = sc["pms"]
        st, inp = O.rollout_4pms(model, gps, pol, x0, T(sc["eps"]), T(sc["masks"]), sc["p_dropout"], T(sc["meas_eps"]),
                                 T(q["std_pos"]), q["pos_idx"], q["vel_idx"], sc["model"]["T"], q["fc"])
    else:
        st, inp = O.rollout(model, gps, pol, x0, T(sc["eps"]), T(sc["masks"]), sc["p_dropout"])
    cost, std = O.expected_cost(oracle_cost(sc, st))
    out = {"states": st.detach().numpy(), "inputs": inp.detach().numpy(), "cost": cost.detach().numpy(), "std_cost": std.numpy()}
    if requires_grad:
        cost.backward()
        out.update(g_log_ls=pol["log_ls"].grad.numpy(), g_centers=pol["centers"].grad.numpy(), g_W=pol["W"].grad.numpy())
        if pol["bias"] is not None:
            out["g_bias"] = pol["bias"].grad.numpy()
    return out

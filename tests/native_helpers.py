"""Scenario -> native (CUDA) objects through the flat operator layer.  Test-only."""
import numpy as np
import torch

from mcpilco_b200 import _ops as ops
from mcpilco_b200 import _pack as P

DEV = "cuda:0"


def G(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device=DEV)


def native_specs(sc):
    return [P.spec_from_dict({"D": sc["D"], "log_ls": g["log_ls"], "lambda": g["lambda"], "mean": g["mean"],
                              "mpk": g["mpk"], "sigma_n": g["sigma_n"]}) for g in sc["gps"]]


def native_fit(sc, golden=None):
    """FittedGp per output: own precompute, or the reference's alpha / K^-1 when `golden` is given."""
    X = G(sc["X"])
    gps = []
    for e, sp in enumerate(native_specs(sc)):
        Linv = None
        if golden is None:  # own precompute: the triangular factor comes along, so forward-only rollouts take the N^2 path
            alpha, Kinv, Linv = ops.gp_precompute(sp, X, G(sc["Y"][:, e:e + 1]), want_Linv=True)
        else:
            alpha, Kinv = G(golden[f"alpha_{e}"]), G(golden[f"Kinv_{e}"])
        gps.append(ops.FittedGp(sp, X, alpha, Kinv, Linv=Linv))
    return gps


def native_model(sc):
    m = sc["model"]
    return P.model_struct(m["kind"], sc["Ds"], sc["Du"], sc["E"], angle=m["angle"], not_angle=m["not_angle"], vel=m["vel"],
                          pos=m["pos"], T=m["T"], use_trig=m["use_trig"])


def native_policy(sc):
    p = sc["policy"]
    Dp = p["centers"].shape[1]
    st = P.policy_struct(p["kind"], p["nb"], Dp, sc["Du"], sc["Ds"], u_max=p["u_max"], scale=p.get("scale"),
                         angle=p.get("angle", ()), non_angle=p.get("non_angle", ()), has_bias=p["bias"] is not None)
    tens = {"log_ls": G(np.log(p["lengthscales"])).reshape(1, -1), "centers": G(p["centers"]), "W": G(p["weight"]),
            "bias": None if p["bias"] is None else G(p["bias"]),
            "target_traj": G(p["target_traj"]) if p["kind"] == "target" else None}
    return st, tens


def native_cost(sc):
    c = sc["cost"]
    if c["kind"] == "cart_pole":
        return P.cost_struct("cart_pole", sc["Ds"], target=c["target"], ls=c["ls"], angle_index=c["angle_index"],
                             pos_index=c["pos_index"]), None
    if c["kind"] == "sat_traj":
        return P.cost_struct("sat_traj", sc["Ds"], ls=c["ls"]), G(c["target_traj"])
    return P.cost_struct(c["kind"], sc["Ds"], target=c["target"], ls=c["ls"], active=c["active"]), None


def native_meas(sc):
    if "pms" not in sc:
        return None
    q = sc["pms"]
    return P.meas_struct(q["pos_idx"], q["vel_idx"], q["std_pos"], q["fc"], sc["model"]["T"])


def native_plan(sc, gps, need_grad=True, fused_cost=True, inject=True, **kw):
    pst, ptens = native_policy(sc)
    cst, ctraj = native_cost(sc) if fused_cost else (None, None)
    noise = dict(eps=G(sc["eps"]), masks=G(sc["masks"]), meas_eps=G(sc["meas_eps"]) if "pms" in sc else None) if inject else {}
    plan = ops.RolloutPlan(native_model(sc), gps, pst, ptens, cost=cst, cost_traj=ctraj, meas=native_meas(sc), M=sc["M"], H=sc["H"],
                           p_dropout=sc["p_dropout"], need_grad=need_grad, **noise, **kw)
    return plan, ptens


def x0_of(sc):
    return G(sc["x0_mean"]).reshape(1, -1) + torch.sqrt(G(sc["x0_var"])).reshape(1, -1) * G(sc["eps0"])

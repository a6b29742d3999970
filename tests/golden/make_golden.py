"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on the seeded
scenarios of tests/scenarios.py with injected noise.  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/<scenario>.npz.  The reference cannot travel to the GPU box, the fixtures do.
Import shims (none touches arithmetic): a stub ``matplotlib`` package, the reference
root on sys.path, and ``gpr_lib.Utils.Parameters_covariance_functions`` imported before any MPK model.
Noise injection: ``_standard_normal`` of torch.distributions.{normal,multivariate_normal}, the
policy's ``f_drop`` attribute and ``torch.randn`` (4PMS) are replaced by feeders that hand out the
scenario's eps / masks in the reference's own draw order.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import scenarios  # noqa: E402

REF = "/root/reference"


def _import_reference():
    shim = tempfile.mkdtemp()
    os.makedirs(os.path.join(shim, "matplotlib"))
    for f in ("__init__.py", "pyplot.py"):
        open(os.path.join(shim, "matplotlib", f), "w").close()
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    import gpr_lib.Utils.Parameters_covariance_functions  # noqa: F401
    import model_learning.Model_learning as ML
    import policy_learning.Cost_function as CF
    import policy_learning.MC_PILCO as MCP
    import policy_learning.Policy as PO
    import gpr_lib.Likelihood.Gaussian_likelihood as LK
    return types.SimpleNamespace(ML=ML, CF=CF, MCP=MCP, PO=PO, LK=LK)


import api_builders as AB  # noqa: E402

CPU = torch.device("cpu")
T = AB.tensor_factory(CPU)


class Feeder:
    """Hands out pre-drawn tensors in call order; checks the requested shape."""

    def __init__(self, items):
        self.items, self.i = list(items), 0

    def __call__(self, shape, *a, **k):
        x = self.items[self.i]; self.i += 1
        assert tuple(shape) == tuple(x.shape), (tuple(shape), tuple(x.shape))
        return x


def nlml_goldens(R, sc, adam_steps=5, lr=0.01):
    """Marginal_log_likelihood value and autograd gradient of every trainable hyper-parameter (Gaussian_likelihood.py:12-24 on
    GP_prior.forward, GP_prior.py:91-115), and the parameters after a few Adam steps of GP_prior.fit_model (GP_prior.py:179-230)."""
    import io, contextlib
    out = {}
    ml = AB.build_model(R, sc, CPU, pretrain=False)
    crit = R.LK.Marginal_log_likelihood()
    for e, gp in enumerate(ml.gp_list):
        X, Y = ml.gp_inputs, ml.gp_output_list[e]
        loss = crit(gp(X), Y)
        loss.backward()
        out[f"nlml_{e}"] = loss.detach().numpy()
        for nm, p in gp.named_parameters():
            if p.requires_grad:
                out[f"nlml_grad_{e}_{nm}"] = p.grad.detach().numpy().copy()
        opt = torch.optim.Adam(gp.parameters(), lr=lr)
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit_model(trainloader=[(X, Y)], optimizer=opt, criterion=crit, N_epoch=adam_steps, N_epoch_print=100)
        for nm, p in gp.named_parameters():
            if p.requires_grad:
                out[f"fit_{e}_{nm}"] = p.detach().numpy().copy()
    return out


def run_scenario(R, name, save=True):
    sc = scenarios.scenario(name)
    out = {}
    ml = AB.build_model(R, sc, CPU)
    rs = np.random.RandomState(7)
    Xs = T(sc["X"][:5] + 0.1 * rs.randn(5, sc["D"]))
    out["Xs"] = Xs.numpy()
    for e, gp in enumerate(ml.gp_list):
        out[f"Kss_{e}"] = gp.get_covariance(Xs, ml.gp_inputs).detach().numpy()
        out[f"Knoise_{e}"] = gp.get_covariance(ml.gp_inputs, flg_noise=True).detach().numpy()
        out[f"kdiag_{e}"] = gp.get_diag_covariance(Xs).detach().numpy()
        out[f"alpha_{e}"] = ml.alpha_list[e].numpy()
        out[f"Kinv_{e}"] = ml.K_X_inv_list[e].numpy()
        mu, var = gp.get_estimate_from_alpha(ml.gp_inputs_tr_list[e], Xs, ml.alpha_list[e], ml.m_X_list[e], ml.K_X_inv_list[e])
        out[f"pmean_{e}"], out[f"pvar_{e}"] = mu.detach().numpy(), var.detach().numpy()
    out.update(nlml_goldens(R, sc))
    if name == "c1":
        thr = 0.5 * torch.sqrt(ml.gp_list[0].get_sigma_n_2())
        with torch.no_grad():
            out["sod_idx_0"] = np.array([int(i) for i in ml.gp_list[0].get_SOD(ml.gp_inputs, ml.gp_output_list[0], thr)])
        out["sod_thr_0"] = thr.detach().numpy()

    # ---- full rollout through the reference's apply_policy ----
    obj = AB.build_pilco(R, sc, ml, CPU, rand_policy=R.PO.Random_exploration)
    pol = obj.control_policy
    M, H, nb, p = sc["M"], sc["H"], sc["policy"]["nb"], sc["p_dropout"]

    feed = Feeder([T(sc["eps0"])] + [T(sc["eps"][t]) for t in range(H - 1)])
    import torch.distributions.multivariate_normal as mvn
    import torch.distributions.normal as nrm
    old = (nrm._standard_normal, mvn._standard_normal, torch.randn)
    nrm._standard_normal = feed
    mvn._standard_normal = feed
    masks = [T(sc["masks"][t]).reshape(M, 1, nb) for t in range(H)]
    cnt = [0]

    def f_drop(x, pp):
        mk = masks[cnt[0]]; cnt[0] += 1
        return x * mk / (1.0 - pp)
    pol.f_drop = f_drop
    if "pms" in sc:
        mfeed = Feeder([T(sc["meas_eps"][t]) for t in range(H - 1)])
        torch.randn = lambda *shape, **k: mfeed(shape)
    try:
        states, inputs = obj.apply_policy(**AB.apply_kwargs(sc, CPU))
    finally:
        nrm._standard_normal, mvn._standard_normal, torch.randn = old
    cost, std_cost = obj.cost_function(states, inputs, 0)
    cost.backward()
    out.update(states=states.detach().numpy(), inputs=inputs.detach().numpy(), cost=cost.detach().numpy(),
               std_cost=std_cost.detach().numpy(), g_log_ls=pol.log_lengthscales.grad.numpy(), g_centers=pol.centers.grad.numpy(),
               g_W=pol.f_linear.weight.grad.numpy())
    if pol.f_linear.bias is not None and pol.f_linear.bias.grad is not None:
        out["g_bias"] = pol.f_linear.bias.grad.numpy()
    # single-step model outputs at t=0 (mean/var of the GP deltas) for the step-level parity test
    with torch.no_grad():
        x0 = T(out["states"][0]); u0 = T(out["inputs"][0])
        nrm._standard_normal = Feeder([T(sc["eps"][0])])
        try:
            nxt, mu, var = ml.get_next_state(x0, u0)
        finally:
            nrm._standard_normal = old[0]
        out.update(step_next=nxt.numpy(), step_mu=mu.numpy(), step_var=var.numpy())
    # remaining cost kinds on the same states (forward values only)
    st = states.detach()
    if name == "delta":
        c = sc["cost"]
        cd, _ = R.CF.Expected_distance(target_state=T(c["target"]), lengthscales=T(c["ls"]), active_dims=c["active"])(st, None)
        out["cost_distance"] = cd.numpy()
    if save:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "cost", float(cost), "std", float(std_cost), "|g_centers|", float(np.abs(out["g_centers"]).max()),
          "min var", min(float(out[f"pvar_{e}"].min()) for e in range(sc["E"])))
    return out


def reinforce_traces(R):
    """Traces of the REFERENCE's reinforce_policy control logic on scripted rollout costs (tests/reinforce_script.py)."""
    import reinforce_script as RS
    out = {}
    for name in RS.SCRIPTS:
        obj = R.MCP.MC_PILCO.__new__(R.MCP.MC_PILCO)  # no simulator / GP needed: apply_policy and cost_function are scripted
        torch.nn.Module.__init__(obj)
        obj.T_sampling, obj.dtype, obj.device, obj.state_dim, obj.input_dim = 0.05, torch.float64, CPU, 1, 1
        for k, v in RS.run(obj, name).items():
            out[f"{name}__{k}"] = v
    np.savez_compressed(os.path.join(HERE, "reinforce_traces.npz"), **out)


if __name__ == "__main__":
    R = _import_reference()
    torch.set_num_threads(1)
    for nm in scenarios.ALL:
        run_scenario(R, nm)
    reinforce_traces(R)

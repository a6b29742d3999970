"""GPU parity through the reference-facing CLASS API: the objects are built by the very code that builds the reference's
objects for the golden vectors (tests/api_builders.py), `apply_policy` / `cost_function` / `cost.backward()` are called the way
`MC_PILCO.reinforce_policy` calls them (reference MC_PILCO.py:484-522), and the results are compared with the golden
vectors the reference produced on the same inputs and injected noise."""
import types

import numpy as np
import pytest
import torch

import api_builders as AB
import helpers as Hh
import scenarios

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def R():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mcpilco_b200.model_learning.Model_learning as ML
    import mcpilco_b200.policy_learning.Cost_function as CF
    import mcpilco_b200.policy_learning.MC_PILCO as MCP
    import mcpilco_b200.policy_learning.Policy as PO
    return types.SimpleNamespace(ML=ML, CF=CF, MCP=MCP, PO=PO)


def relmax(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


def noise_of(sc):
    T = AB.tensor_factory(DEV)
    nz = dict(eps0=T(sc["eps0"]), eps=T(sc["eps"]), masks=T(sc["masks"]))
    if "pms" in sc:
        nz["meas_eps"] = T(sc["meas_eps"])
    return nz


@pytest.mark.parametrize("name", scenarios.ALL)
def test_gp_objects(R, name):
    """gpr_lib surface: get_covariance / get_diag_covariance / get_estimate_from_alpha / pretrain_gp state."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    T = AB.tensor_factory(DEV)
    ml = AB.build_model(R, sc, DEV)
    Xs = T(g["Xs"])
    for e, gp in enumerate(ml.gp_list):
        assert relmax(gp.get_covariance(Xs, ml.gp_inputs), g[f"Kss_{e}"]) < 1e-12
        assert relmax(gp.get_covariance(ml.gp_inputs, flg_noise=True), g[f"Knoise_{e}"]) < 1e-12
        assert relmax(gp.get_diag_covariance(Xs), g[f"kdiag_{e}"]) < 1e-12
        assert relmax(ml.alpha_list[e], g[f"alpha_{e}"]) < 1e-6 and ml.alpha_list[e].shape == g[f"alpha_{e}"].shape
        assert relmax(ml.K_X_inv_list[e], g[f"Kinv_{e}"]) < 1e-6
        assert ml.m_X_list[e].shape == (sc["N"], 1) and ml.gp_inputs_tr_list[e].shape == (sc["N"], sc["D"])
        mu, var = gp.get_estimate_from_alpha(ml.gp_inputs_tr_list[e], Xs, T(g[f"alpha_{e}"]), ml.m_X_list[e], T(g[f"Kinv_{e}"]))
        assert mu.shape == g[f"pmean_{e}"].shape and var.shape == g[f"pvar_{e}"].shape
        assert relmax(mu, g[f"pmean_{e}"]) < 1e-10 and relmax(var, g[f"pvar_{e}"]) < 1e-8
    m_X, K_X, K_X_inv, log_det = ml.gp_list[0](ml.gp_inputs)
    ref_logdet = np.linalg.slogdet(g["Knoise_0"])[1]
    assert abs(float(log_det) - ref_logdet) < 1e-8 * abs(ref_logdet) and relmax(K_X, g["Knoise_0"]) < 1e-12


@pytest.mark.parametrize("name", scenarios.ALL)
def test_one_step(R, name):
    """Model_learning.get_next_state with the reparameterisation noise injected through torch.distributions."""
    import torch.distributions.normal as nrm
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    T = AB.tensor_factory(DEV)
    ml = AB.build_model(R, sc, DEV)
    old = nrm._standard_normal
    nrm._standard_normal = lambda shape, dtype, device: T(sc["eps"][0])
    try:
        with torch.no_grad():
            nxt, mu, var = ml.get_next_state(T(g["states"][0]), T(g["inputs"][0]))
    finally:
        nrm._standard_normal = old
    assert relmax(mu, g["step_mu"]) < 1e-5 and relmax(var, g["step_var"]) < 1e-5 and relmax(nxt, g["step_next"]) < 1e-5


def test_sod_selection_and_sod_model(R):
    sc, g = scenarios.scenario("c1"), Hh.load_golden("c1")
    ml = AB.build_model(R, sc, DEV, pretrain=False)
    with torch.no_grad():
        idx = ml.gp_list[0].get_SOD(ml.gp_inputs, ml.gp_output_list[0], torch.tensor(g["sod_thr_0"], device=DEV))
    assert [int(i) for i in idx] == [int(i) for i in g["sod_idx_0"]]
    ml2 = AB.build_model(R, sc, DEV, approximation={"SOD_threshold_mode": "relative", "SOD_threshold": 0.5, "flg_SOD_permutation": False})
    assert [int(i) for i in ml2.SOD_indices[0]] == [int(i) for i in g["sod_idx_0"]]
    assert ml2.gp_inputs_tr_list[0].shape[0] == len(g["sod_idx_0"]) and ml2.K_X_inv_list[0].shape[0] == len(g["sod_idx_0"])
    # per-output training sets of different sizes go through the fused rollout
    obj = AB.build_pilco(R, sc, ml2, DEV)
    st, inp = obj.apply_policy(**AB.apply_kwargs(sc, DEV), _noise=noise_of(sc))
    assert torch.isfinite(st).all() and st.shape == (sc["H"], sc["M"], sc["Ds"])


@pytest.mark.parametrize("name", scenarios.ALL)
def test_policy_module(R, name):
    """Stand-alone policy call on a batch of states (dropout off) against the golden inputs' formula via the oracle."""
    from oracle import mcpilco_oracle as O
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    pol = AB.build_policy(R, sc, DEV)
    x = AB.tensor_factory(DEV)(g["states"][1])
    u = pol(x, t=1, p_dropout=0.0)
    ref = O.policy_apply(Hh.oracle_policy(sc), Hh.T(g["states"][1]), 1, None, 0.0)
    assert u.shape == ref.shape and relmax(u, ref.numpy()) < 1e-10
    sd = pol.state_dict()
    assert {"log_lengthscales", "centers", "f_linear.weight"} <= set(sd)


@pytest.mark.parametrize("name", scenarios.ALL)
def test_apply_policy_cost_backward(R, name):
    """The reinforce_policy inner sequence: apply_policy -> cost_function -> cost.backward() -> .grad on the policy tensors."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    pol = obj.control_policy
    states, inputs = obj.apply_policy(**AB.apply_kwargs(sc, DEV), _noise=noise_of(sc))
    cost, std_cost = obj.cost_function(states, inputs, 0)
    assert cost.dim() == 0 and std_cost.dim() == 0 and states.shape == g["states"].shape and inputs.shape == g["inputs"].shape
    assert relmax(states, g["states"]) < 1e-5 and relmax(inputs, g["inputs"]) < 1e-5
    assert abs(float(cost.detach()) - float(g["cost"])) < 1e-5 * abs(float(g["cost"]))
    assert abs(float(std_cost) - float(g["std_cost"])) < 1e-4 * abs(float(g["std_cost"]))
    cost.backward()
    assert relmax(pol.log_lengthscales.grad, g["g_log_ls"]) < 1e-4 and pol.log_lengthscales.grad.shape == g["g_log_ls"].shape
    assert relmax(pol.centers.grad, g["g_centers"]) < 1e-4
    assert relmax(pol.f_linear.weight.grad, g["g_W"]) < 1e-4
    if "g_bias" in g and pol.f_linear.bias.requires_grad:
        assert relmax(pol.f_linear.bias.grad, g["g_bias"]) < 1e-4
    # gradients accumulate like autograd's: a second identical pass doubles .grad
    states, inputs = obj.apply_policy(**AB.apply_kwargs(sc, DEV), _noise=noise_of(sc))
    obj.cost_function(states, inputs, 0)[0].backward()
    assert relmax(pol.centers.grad, 2 * g["g_centers"]) < 1e-4


@pytest.mark.parametrize("name", ["c1", "c4"])
def test_user_cost_lambda_generic_path(R, name):
    """Expected_cost(cost_function=<user lambda>): torch ops on the device, autograd hands grad_states to the CUDA backward."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    cls, par = AB.cost_spec(R, sc, DEV)
    concrete = cls(**par)
    obj.cost_function = R.CF.Expected_cost(concrete.cost_function)  # same maths, but opaque to the fused path
    states, inputs = obj.apply_policy(**AB.apply_kwargs(sc, DEV), _noise=noise_of(sc))
    cost, std_cost = obj.cost_function(states, inputs, 0)
    assert abs(float(cost.detach()) - float(g["cost"])) < 1e-5 * abs(float(g["cost"]))
    cost.backward()
    assert relmax(obj.control_policy.centers.grad, g["g_centers"]) < 1e-4
    assert relmax(obj.control_policy.f_linear.weight.grad, g["g_W"]) < 1e-4


def test_no_grad_rollout_and_dropout_free(R):
    sc, g = scenarios.scenario("c2"), Hh.load_golden("c2")
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    with torch.no_grad():
        st, inp = obj.apply_policy(**AB.apply_kwargs(sc, DEV), _noise=noise_of(sc))
        c, s = obj.cost_function(st, inp, 0)
    assert not st.requires_grad and not c.requires_grad
    assert abs(float(c) - float(g["cost"])) < 1e-5 * abs(float(g["cost"]))
    kw = AB.apply_kwargs(sc, DEV); kw["p_dropout"] = 0.0
    outs = []
    for seed in (3, 3, 4):
        fresh = AB.build_pilco(R, sc, ml, DEV)
        torch.manual_seed(seed)
        with torch.no_grad():
            outs.append(fresh.apply_policy(**kw)[0])
    assert torch.equal(outs[0], outs[1])      # same torch seed -> same Philox key -> same particles
    assert not torch.equal(outs[0], outs[2])


def test_reinforce_policy_loop(R):
    """A short optimisation run through reinforce_policy: cost decreases, outputs have the reference's types/shapes."""
    sc = scenarios.scenario("c2")
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    T = AB.tensor_factory(DEV)
    torch.manual_seed(0)
    out = obj.reinforce_policy(T_control=sc["H"] * obj.T_sampling + 1e-9, num_particles=64, trial_index=0,
                               particles_initial_state_mean=T(sc["x0_mean"]), particles_initial_state_var=T(sc["x0_var"]),
                               flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None,
                               flg_particles_init_multi_gauss=False, opt_steps_list=[30], lr_list=[0.05],
                               f_optimizer="lambda p, lr : torch.optim.Adam(p, lr)", num_step_print=10, p_dropout_list=[0.1],
                               policy_reinit_dict=dict(lenghtscales_par=sc["policy"]["lengthscales"], centers_par=np.ones((20, 5)),
                                                       weight_par=1.0))
    cost_list, std_list, states, inputs = out
    assert cost_list.shape == (30,) and std_list.shape == (30,) and isinstance(states, np.ndarray)
    assert states.shape == (sc["H"], 64, 4) and inputs.shape == (sc["H"], 64, 1)
    assert cost_list[-5:].mean() < cost_list[:5].mean()


@pytest.mark.parametrize("name", scenarios.ALL)
def test_marginal_likelihood_and_gradients(R, name):
    """Marginal_log_likelihood()(gp(X), Y).backward(): value and every hyper-parameter gradient against the reference's autograd."""
    import mcpilco_b200.gpr_lib.Likelihood.Gaussian_likelihood as LK
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    ml = AB.build_model(R, sc, DEV, pretrain=False)
    crit = LK.Marginal_log_likelihood()
    for e, gp in enumerate(ml.gp_list):
        loss = crit(gp(ml.gp_inputs), ml.gp_output_list[e])
        assert loss.shape == (1, 1)
        assert abs(float(loss.detach()) - float(g[f"nlml_{e}"])) < 1e-9 * abs(float(g[f"nlml_{e}"]))
        loss.backward()
        names = [nm for nm, p in gp.named_parameters() if p.requires_grad]
        assert set(f"nlml_grad_{e}_{nm}" for nm in names) == set(k for k in g if k.startswith(f"nlml_grad_{e}_"))
        for nm, p in gp.named_parameters():
            if p.requires_grad:
                ref = g[f"nlml_grad_{e}_{nm}"]
                assert p.grad.shape == ref.shape and relmax(p.grad, ref) < 1e-6, (nm, relmax(p.grad, ref))


@pytest.mark.parametrize("name", ["c1", "c2", "delta"])
def test_fit_model_adam_steps(R, name):
    """Five Adam epochs of GP_prior.fit_model land on the reference's parameters."""
    import contextlib, io
    import mcpilco_b200.gpr_lib.Likelihood.Gaussian_likelihood as LK
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    ml = AB.build_model(R, sc, DEV, pretrain=False)
    for e, gp in enumerate(ml.gp_list):
        opt = torch.optim.Adam(gp.parameters(), lr=0.01)
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit_model(trainloader=[(ml.gp_inputs, ml.gp_output_list[e])], optimizer=opt, criterion=LK.Marginal_log_likelihood(), N_epoch=5,
                         N_epoch_print=100)
        for nm, p in gp.named_parameters():
            if p.requires_grad:
                np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"fit_{e}_{nm}"], rtol=1e-6, atol=1e-7)


def test_reinforce_model_trains_and_pretrains(R):
    """Model_learning.reinforce_model with the reference's option dict: likelihood goes down, the model is ready for rollouts."""
    import contextlib, io
    import mcpilco_b200.gpr_lib.Likelihood.Gaussian_likelihood as LK
    sc = scenarios.scenario("c2")
    ml = AB.build_model(R, sc, DEV, pretrain=False)
    crit = LK.Marginal_log_likelihood()
    before = [float(crit(gp(ml.gp_inputs), ml.gp_output_list[e]).detach()) for e, gp in enumerate(ml.gp_list)]
    opt = {"f_optimizer": "lambda p : torch.optim.Adam(p, lr=0.01)", "criterion": LK.Marginal_log_likelihood, "N_epoch": 40, "N_epoch_print": 100}
    with contextlib.redirect_stdout(io.StringIO()):
        ml.reinforce_model(optimization_opt_list=[opt] * sc["E"])
    after = [float(crit(gp(ml.gp_inputs), ml.gp_output_list[e]).detach()) for e, gp in enumerate(ml.gp_list)]
    assert all(a < b for a, b in zip(after, before))
    assert all(k is not None for k in ml.K_X_inv_list) and ml.alpha_list[0].shape == (sc["N"], 1)


def test_initial_particle_distributions(R):
    """Gaussian / uniform / multi-modal Gaussian initial clouds of apply_policy (reference MC_PILCO.py:635-657): moments of the
    Philox-generated particles, and states[0] is the cloud."""
    sc = scenarios.scenario("c2")
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    T = AB.tensor_factory(DEV)
    M = 200000
    base = dict(flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None,
                flg_particles_init_multi_gauss=False, num_particles=M, T_control=1, p_dropout=0.0)
    mean, var = T([0.5, -1.0, 2.0, 0.0]), T([1e-2, 4.0, 1e-4, 0.25])
    with torch.no_grad():
        st, _ = obj.apply_policy(particles_initial_state_mean=mean, particles_initial_state_var=var, **base)
    x0 = st[0]
    assert x0.shape == (M, 4)
    assert torch.allclose(x0.mean(0), mean, atol=5 * float(var.max().sqrt()) / M ** 0.5)
    assert torch.allclose(x0.var(0), var, rtol=0.03)
    z = (x0 - mean) / var.sqrt()
    assert abs(float((z ** 3).mean())) < 0.03 and abs(float((z ** 4).mean()) - 3.0) < 0.1      # skewness 0, kurtosis 3
    assert abs(float(torch.corrcoef(z.t())[0, 1])) < 0.01                                   # independent dimensions
    lo, up = T([-1.0, 0.0, 2.0, -3.0]), T([1.0, 0.5, 2.5, 3.0])
    kw = dict(base); kw.update(flg_particles_init_uniform=True, particles_init_up_bound=up, particles_init_low_bound=lo)
    with torch.no_grad():
        st, _ = obj.apply_policy(particles_initial_state_mean=mean, particles_initial_state_var=var, **kw)
    x0 = st[0]
    assert bool((x0 >= lo).all()) and bool((x0 <= up).all())
    assert torch.allclose(x0.mean(0), (lo + up) / 2, atol=0.02) and torch.allclose(x0.var(0), (up - lo) ** 2 / 12, rtol=0.03)
    means = T([[0.0, 0.0, 0.0, 0.0], [10.0, 10.0, 10.0, 10.0], [-10.0, 5.0, 0.0, 1.0]])
    vars_ = T([[1e-2] * 4, [1e-4] * 4, [1.0] * 4])
    kw = dict(base); kw.update(flg_particles_init_multi_gauss=True)
    with torch.no_grad():
        st, _ = obj.apply_policy(particles_initial_state_mean=means, particles_initial_state_var=vars_, **kw)
    x0 = st[0]
    mode = torch.cdist(x0, means).argmin(1)
    frac = torch.bincount(mode, minlength=3).double() / M
    assert torch.allclose(frac, torch.full((3,), 1 / 3, dtype=torch.float64, device=DEV), atol=0.01)
    for k in range(3):
        sel = x0[mode == k]
        assert torch.allclose(sel.mean(0), means[k], atol=0.02) and torch.allclose(sel.var(0), vars_[k], rtol=0.05)


def test_mean_rollout_along_recorded_inputs(R):
    """MC_PILCO.rollout (reference MC_PILCO.py:347-373): one particle, particle_pred=False, recorded input trajectory — against the
    oracle's next_state iterated with the same inputs."""
    from oracle import mcpilco_oracle as O
    sc, g = scenarios.scenario("c1"), Hh.load_golden("c1")
    ml = AB.build_model(R, sc, DEV)
    obj = AB.build_pilco(R, sc, ml, DEV)
    Tn = 6
    obj.state_samples_history = [g["states"][:Tn, 0, :]]
    obj.input_samples_history = [g["inputs"][:Tn, 0, :]]
    traj = obj.rollout(0)
    assert traj.shape == (Tn, 4)
    X = Hh.T(sc["X"])
    gps = [(sp, X, Hh.T(g[f"alpha_{e}"]), Hh.T(g[f"Kinv_{e}"])) for e, sp in enumerate(Hh.oracle_specs(sc))]
    x = Hh.T(g["states"][0:1, 0, :])
    ref = [x]
    for t in range(1, Tn):
        x, _, _ = O.next_state(Hh.oracle_model(sc), gps, x, Hh.T(g["inputs"][t - 1:t, 0, :]), None, particle_pred=False)
        ref.append(x)
    np.testing.assert_allclose(traj, torch.cat(ref).numpy(), rtol=1e-6, atol=1e-9)


def test_example_trial_loop_runs():
    """examples/cartpole_swingup.py: data -> GP training -> policy optimisation -> apply, twice, all through the mirrored class API."""
    import os, subprocess, sys
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "cartpole_swingup.py"), "--trials", "2", "--opt-steps", "40", "--gp-epochs", "60"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("trial ")]
    assert len(lines) == 2
    import re
    c0, c1 = map(float, re.search(r"particle cost ([0-9.]+) -> ([0-9.]+)", lines[0]).groups())
    assert c1 < c0

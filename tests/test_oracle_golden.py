"""CPU: pins oracle/mcpilco_oracle.py against golden vectors produced by the real reference
(tests/golden/make_golden.py).  Tolerances are rounding-level: same formulas, same op order."""
import numpy as np
import pytest
import torch

import helpers as Hh
import scenarios
from oracle import mcpilco_oracle as O

T = Hh.T


def close(a, b, rtol=1e-10, atol=1e-13):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", scenarios.ALL)
def test_kernel_and_fit(name):
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X, Xs = T(sc["X"]), T(g["Xs"])
    for e, sp in enumerate(Hh.oracle_specs(sc)):
        close(O.gp_cov(sp, Xs, X), g[f"Kss_{e}"], 1e-12)
        close(O.gp_cov(sp, X, None, noise=True), g[f"Knoise_{e}"], 1e-12)
        close(O.gp_diag(sp, Xs), g[f"kdiag_{e}"], 1e-12)
        alpha, _, Kinv = O.gp_fit(sp, X, T(sc["Y"][:, e:e + 1]))
        # conditioning-limited: cond(K) ~ 1e5..1e6 here
        close(Kinv, g[f"Kinv_{e}"], 1e-7, 1e-7 * np.abs(g[f"Kinv_{e}"]).max())
        close(alpha, g[f"alpha_{e}"], 1e-7, 1e-7 * np.abs(g[f"alpha_{e}"]).max())
        # posterior from the reference's own alpha / K^-1: formula parity at rounding level
        mu, var = O.gp_predict(sp, X, T(g[f"alpha_{e}"]), T(g[f"Kinv_{e}"]), Xs)
        close(mu, g[f"pmean_{e}"], 1e-11, 1e-13)
        close(var, g[f"pvar_{e}"], 1e-9, 1e-13)


def test_sod_selection():
    sc, g = scenarios.scenario("c1"), Hh.load_golden("c1")
    sp = Hh.oracle_specs(sc)[0]
    idx = O.sod_select(sp, T(sc["X"]), T(sc["Y"][:, 0:1]), T(g["sod_thr_0"]))
    assert idx == list(g["sod_idx_0"])


@pytest.mark.parametrize("name", scenarios.ALL)
def test_step_and_rollout(name):
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X = T(sc["X"])
    gps = [(sp, X, T(g[f"alpha_{e}"]), T(g[f"Kinv_{e}"])) for e, sp in enumerate(Hh.oracle_specs(sc))]
    nxt, mu, var = O.next_state(Hh.oracle_model(sc), gps, T(g["states"][0]), T(g["inputs"][0]), T(sc["eps"][0]))
    close(mu, g["step_mu"], 1e-11); close(var, g["step_var"], 1e-9); close(nxt, g["step_next"], 1e-11)
    out = Hh.oracle_rollout(sc, gps)
    close(out["states"], g["states"], 1e-9, 1e-12)
    close(out["inputs"], g["inputs"], 1e-9, 1e-12)
    close(out["cost"], g["cost"], 1e-11); close(out["std_cost"], g["std_cost"], 1e-10)
    for k in ("g_log_ls", "g_centers", "g_W", "g_bias"):
        if k in g:
            close(out[k], g[k], 1e-8, 1e-9 * np.abs(g[k]).max())


def test_other_costs():
    sc, g = scenarios.scenario("delta"), Hh.load_golden("delta")
    c = sc["cost"]
    cd, _ = O.expected_cost(O.cost_distance(T(g["states"]), T(c["target"]), T(c["ls"]), c["active"]))
    close(cd, g["cost_distance"], 1e-12)


def test_butter1_matches_scipy():
    from scipy import signal
    for fc in (0.5, 0.2, 0.8):
        b, a = signal.butter(1, fc)
        bo, ao = O.butter1(fc)
        close(bo, b, 1e-12, 1e-15); close(ao, a, 1e-12, 1e-15)


@pytest.mark.parametrize("name", scenarios.ALL)
def test_nlml_and_gradient(name):
    """Training objective and its hyper-parameter gradients (log-space parameters, as the reference stores them)."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X = T(sc["X"])
    for e, sp in enumerate(Hh.oracle_specs(sc)):
        leaves = [sp["se"]["log_ls"].requires_grad_(True), sp["sigma_n_log"].requires_grad_(True)] + [m["log_par"].requires_grad_(True) for m in sp["mpk"]]
        loss = O.nlml(sp, X, T(sc["Y"][:, e:e + 1]))
        close(loss.detach(), g[f"nlml_{e}"], 1e-10)
        grads = torch.autograd.grad(loss.sum(), leaves)
        pre = "gp_list.0." if sp["mpk"] else ""
        close(grads[0], g[f"nlml_grad_{e}_{pre}log_lengthscales_par"], 1e-7, 1e-9)
        close(grads[1], g[f"nlml_grad_{e}_{pre}sigma_n_log"].reshape(()), 1e-7, 1e-9)
        for k in range(len(sp["mpk"])):
            close(grads[2 + k], g[f"nlml_grad_{e}_gp_list.1.gp_list.{k}.Sigma_pos_par"], 1e-7, 1e-9)

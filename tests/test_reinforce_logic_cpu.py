"""CPU: control logic of MC_PILCO.reinforce_policy against traces recorded from the REFERENCE's reinforce_policy
(reference MC_PILCO.py:375-613) driven by the same scripted rollout costs (tests/reinforce_script.py): NaN re-sampling, policy
re-initialisation and restart, cost-difference monitors, learning-rate halving with dropout reduction, early exit.  No GP, no CUDA:
apply_policy / cost_function are replaced by the script on both sides."""
import os

import numpy as np
import pytest
import torch

import reinforce_script as RS

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reinforce_traces.npz")))


def make_obj():
    import mcpilco_b200.policy_learning.MC_PILCO as MCP
    obj = MCP.MC_PILCO.__new__(MCP.MC_PILCO)
    torch.nn.Module.__init__(obj)
    obj.T_sampling, obj.dtype, obj.device, obj.state_dim, obj.input_dim = 0.05, torch.float64, torch.device("cpu"), 1, 1
    obj._trial_index, obj._seed_base, obj._rollouts = None, None, 0
    return obj


@pytest.mark.parametrize("name", RS.SCRIPTS)
def test_reinforce_policy_control_logic(name):
    res = RS.run(make_obj(), name)
    g = {k.split("__", 1)[1]: v for k, v in GOLD.items() if k.startswith(name + "__")}
    assert len(res["cost_list"]) == len(g["cost_list"])                      # same number of optimisation steps (early exit included)
    assert int(res["n_rollouts"]) == int(g["n_rollouts"])                    # same number of rollouts (NaN re-sampling included)
    assert int(res["reinits"]) == int(g["reinits"])                          # same number of policy re-initialisations
    np.testing.assert_array_equal(res["dropouts"], g["dropouts"])            # same p_dropout handed to every rollout
    np.testing.assert_allclose(res["cost_list"], g["cost_list"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(res["std_list"], g["std_list"], rtol=1e-12)
    np.testing.assert_allclose(res["w_final"], g["w_final"], rtol=1e-10)     # same optimiser trajectory (lr schedule included)
    assert res["states"].shape == g["states"].shape


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line (library banners and prints go to stderr); the reference arm is the oracle port."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--train-points", "128", "--particles-per-gpu", "32", "--horizon", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0

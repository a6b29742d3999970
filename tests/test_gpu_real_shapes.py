"""GPU parity at the reference's REAL configuration sizes (BASELINE.json configs 1-4, SURVEY.md §8: C1-C4) and at a mid-size
point of the sweep (C5) that runs the bench's own kernels with the bench's policy size.

Sizes: C1/C2  N = 300, M = 400, H = 60, nb = 200 (test_mcpilco_cartpole.py:124,199, test_mcpilco_cartpole_rbf_ker.py);
C3  H = 90 with the 4PMS measurement model (test_mcpilco4pms_cartpole.py:104,171); C4  N = 400, M = 200, H = 200, nb = 400, D = 24,
E = 6 (test_mcpilco_ur5_mujoco.py:127,195); C1 at its first trial, N = 60.  nb = 200 / 400 means 7 / 13 warps in the backward kernel
(cross-warp reductions, strided basis loops) — code the miniature scenarios of test_gpu_parity.py never reach.

The CUDA rollout (forward + hand-written backward) is compared with the CPU oracle on the same seeded data and the same injected
noise (initial particles, reparameterisation noise, dropout masks, measurement noise), with each side's OWN precompute, on every
code path the shape can take.  Tolerances are the north star's: trajectories and cost 1e-5 relative, policy gradients 1e-4."""
import numpy as np
import pytest
import torch

import helpers as Hh

pytestmark = pytest.mark.gpu

REL_VAL, REL_GRAD = 1e-5, 1e-4
_oracle_cache = {}


@pytest.fixture(scope="module")
def nh():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import native_helpers
    return native_helpers


def relmax(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


def oracle_of(key, sc):
    if key not in _oracle_cache:
        torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
        _oracle_cache[key] = Hh.oracle_rollout(sc)
    return _oracle_cache[key]


PATHS = {"persistent": {"MCPILCO_PERSIST": "1"}, "fused-small": {"MCPILCO_NO_PERSIST": "1"}, "per-step": {"MCPILCO_NO_SMALL_PATH": "1"},
         "per-output-chains": {"MCPILCO_NO_SMALL_PATH": "1", "MCPILCO_NO_BATCHED_STEP": "1"}, "no-pdl": {"MCPILCO_NO_PERSIST": "1", "MCPILCO_NO_PDL": "1"}}
CASES = [("c1", "persistent"), ("c1", "fused-small"), ("c1", "per-step"), ("c1", "no-pdl"), ("c2", "persistent"), ("c2", "fused-small"),
         ("c2", "per-step"), ("c3", "persistent"), ("c3", "fused-small"), ("c3", "per-step"), ("c4", "per-step"), ("c4", "per-output-chains"),
         ("c1_first_trial", "persistent"), ("c1_first_trial", "fused-small"), ("c1_first_trial", "per-step")]


def check(nh, sc, ref, tag):
    gps = nh.native_fit(sc)
    plan, _ = nh.native_plan(sc, gps, need_grad=True)
    states, inputs = plan.forward(nh.x0_of(sc))
    gr = plan.backward(grad_cost=1.0)
    torch.cuda.synchronize()
    errs = {"states": relmax(states, ref["states"]), "inputs": relmax(inputs, ref["inputs"]),
            "cost": abs(float(plan.cost_out[0]) - float(ref["cost"])) / abs(float(ref["cost"])),
            "std_cost": abs(float(plan.cost_out[1]) - float(ref["std_cost"])) / abs(float(ref["std_cost"]))}
    for k in ("log_ls", "centers", "W"):
        errs["g_" + k] = relmax(gr[k], ref["g_" + k])
    print(tag, {k: "%.2e" % v for k, v in errs.items()})
    assert not torch.isnan(states).any()
    assert errs["states"] < REL_VAL and errs["inputs"] < REL_VAL and errs["cost"] < REL_VAL and errs["std_cost"] < 1e-4, errs
    assert max(errs["g_log_ls"], errs["g_centers"], errs["g_W"]) < REL_GRAD, errs
    return errs


@pytest.mark.parametrize("key,path", CASES)
def test_real_shape_rollout_vs_oracle(nh, monkeypatch, key, path):
    from mcpilco_b200 import workloads as W
    for v in ("MCPILCO_NO_SMALL_PATH", "MCPILCO_NO_BATCHED_STEP", "MCPILCO_NO_PDL", "MCPILCO_NO_PERSIST", "MCPILCO_PERSIST"):
        monkeypatch.delenv(v, raising=False)
    for k, v in PATHS[path].items():
        monkeypatch.setenv(k, v)
    sc = W.real_shape(key)
    assert (sc["N"], sc["M"], sc["H"], sc["policy"]["nb"]) == W.REAL_SHAPES[key][1:]
    check(nh, sc, oracle_of(key, sc), "%s/%s" % (key, path))


@pytest.mark.parametrize("nb", [7, 32, 33, 224, 225, 512])
@pytest.mark.parametrize("key", ["c1", "c3"])
def test_backward_sweep_basis_function_counts_vs_oracle(nh, monkeypatch, key, nb):
    """The backward sweep puts one thread per basis function plus one chain warp: basis-function counts below / at / above a warp, at
    the 256-thread variant's limit (224 + 32), just above it and at the maximum (16 basis warps + the chain warp), with and without
    the measurement model.  Cart-pole shape with few particles; forward + backward against the oracle."""
    from mcpilco_b200 import workloads as W
    for v in ("MCPILCO_NO_SMALL_PATH", "MCPILCO_NO_BATCHED_STEP", "MCPILCO_NO_PDL", "MCPILCO_NO_PERSIST", "MCPILCO_PERSIST"):
        monkeypatch.delenv(v, raising=False)
    sc = W.real_shape(key, N=64, M=37, H=9, nb=nb)
    check(nh, sc, oracle_of("bwd-%s-%d" % (key, nb), sc), "%s/nb=%d" % (key, nb))


@pytest.mark.parametrize("key,M", [("c1", 1000), ("c3", 433), ("c2", 2048)])
def test_cluster_kernel_several_batches_per_cluster_vs_oracle(nh, monkeypatch, key, M):
    """More particles than the 15 co-resident clusters x 27 take in one pass: every cluster walks several batches (state reset, the
    K^-1 slice ring carried across batches), the last batch ragged; 2048 is the path's upper limit."""
    from mcpilco_b200 import workloads as W
    for v in ("MCPILCO_NO_SMALL_PATH", "MCPILCO_NO_BATCHED_STEP", "MCPILCO_NO_PDL", "MCPILCO_NO_PERSIST"):
        monkeypatch.delenv(v, raising=False)
    monkeypatch.setenv("MCPILCO_PERSIST", "1")
    sc = W.real_shape(key, N=90, M=M, H=5)
    check(nh, sc, oracle_of("batches-%s-%d" % (key, M), sc), "%s/M=%d" % (key, M))


def test_sweep_midsize_bench_kernels_vs_oracle(nh):
    """C5 at N = 2048, M = 4096, H = 4 with the bench's policy (nb = 200): more than 2048 particles, so this is the per-step path the
    bench times — cov_fast, the TMA-pipelined DMMA contraction over several 128-row tiles, the fast reduce, the 7-warp backward."""
    from mcpilco_b200 import workloads as W
    sc = W.real_shape("c1", N=2048, M=4096, H=4)
    for g in sc["gps"]:
        g["sigma_n"] = 0.1
    check(nh, sc, oracle_of("mid", sc), "c5-mid")


def test_real_shape_forward_only_matches_forward_with_grad(nh):
    """The no-grad rollout (reference MC_PILCO.py:430-456) takes a cheaper contraction; trajectories and cost must agree with the
    differentiable rollout's to 1e-8 (same noise)."""
    from mcpilco_b200 import workloads as W
    for key, env in (("c1", None), ("c4", None)):
        sc = W.real_shape(key)
        gps = nh.native_fit(sc)
        out = []
        for need_grad in (True, False):
            plan, _ = nh.native_plan(sc, gps, need_grad=need_grad)
            s, u = plan.forward(nh.x0_of(sc))
            out.append((s.clone(), u.clone(), plan.cost_out.clone()))
        assert relmax(out[1][0], out[0][0].cpu().numpy()) < 1e-8 and relmax(out[1][1], out[0][1].cpu().numpy()) < 1e-8
        assert relmax(out[1][2], out[0][2].cpu().numpy()) < 1e-8


# ---------------------------------------------------------------------------------------------------------------------
# sharded gradients: the autograd node's shard weighting and flat all-reduce, emulated on one GPU
# ---------------------------------------------------------------------------------------------------------------------
def _build_obj(sc):
    import types
    import api_builders as AB
    import mcpilco_b200.model_learning.Model_learning as ML
    import mcpilco_b200.policy_learning.Cost_function as CF
    import mcpilco_b200.policy_learning.MC_PILCO as MCP
    import mcpilco_b200.policy_learning.Policy as PO
    R = types.SimpleNamespace(ML=ML, CF=CF, MCP=MCP, PO=PO)
    dev = torch.device("cuda:0")
    return AB, AB.build_pilco(R, sc, AB.build_model(R, sc, dev), dev), dev


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("generic_cost", [False, True])
def test_sharded_gradients_equal_unsharded(nh, monkeypatch, world, generic_cost):
    """`world` ranks emulated one after the other on one GPU with the collectives replaced by a recorder (pass 1: every rank runs its
    shard and records what it would contribute) and a replayer (pass 2: every rank runs again and receives the recorded all-gather /
    all-reduce results).  Philox keys are a function of (seed base, rollout counter, global particle id), so the passes draw the same
    noise.  Every rank must end with the unsharded run's cost and .grad (<= 1e-13 relative: the particle sums are split
    differently).  generic_cost: a user cost lambda instead of the fused cost (states gradients enter through grad_states)."""
    import scenarios
    from mcpilco_b200 import distributed as D
    sc = dict(scenarios.scenario("c1"))
    sc["M"] = 401                                  # uneven shards
    AB, obj, dev = _build_obj(sc)
    if generic_cost:
        import mcpilco_b200.policy_learning.Cost_function as CF
        tgt = torch.tensor([0.3, 0.0, 3.0, 0.0], dtype=torch.float64, device=dev)
        obj.cost_function = CF.Expected_cost(lambda states, inputs, trial_index: 1 - torch.exp(-((states - tgt) ** 2).sum(-1) / 9.0))
    kw = AB.apply_kwargs(sc, dev)
    kw["num_particles"] = sc["M"]
    pol = obj.control_policy
    params = [pol.log_lengthscales, pol.centers, pol.f_linear.weight]

    def run():
        obj._seed_base, obj._rollouts = 1234, 0
        for p in params:
            p.grad = None
        states, inputs = obj.apply_policy(**kw)
        cost, std = obj.cost_function(states, inputs, 0)
        cost.backward()
        return cost.detach().clone(), std.detach().clone(), [p.grad.detach().clone() for p in params], states.shape[1]

    cost0, std0, g0, m0 = run()
    assert m0 == sc["M"]
    rec = {"stats": {}, "flat": {}, "scalars": {}}
    state = {"rank": 0, "replay": False}
    monkeypatch.setattr(D, "world", lambda: (state["rank"], world, "fake-group"))

    def fake_gather(local, group, ws):
        if not state["replay"]:
            rec["stats"][state["rank"]] = local.clone()
            return torch.stack([local] * ws)
        return torch.stack([rec["stats"][r] for r in range(ws)])

    def fake_allreduce(flat, group):
        key = flat.numel()
        if not state["replay"]:
            rec["flat"].setdefault(key, {})[state["rank"]] = flat.clone()
            return flat
        tot = rec["flat"][key][0].clone()
        for r in range(1, world):
            tot = tot + rec["flat"][key][r]
        flat.copy_(tot)
        return flat

    monkeypatch.setattr(D, "gather_cost_stats", fake_gather)
    monkeypatch.setattr(D, "allreduce_sum_", fake_allreduce)
    counts = []
    for state["replay"] in (False, True):
        for r in range(world):
            state["rank"] = r
            cost, std, g, m = run()
            if state["replay"]:
                counts.append(m)
                assert float(cost) == pytest.approx(float(cost0), rel=1e-14)
                assert float(std) == pytest.approx(float(std0), rel=1e-11)
                for a, b in zip(g, g0):
                    assert float((a - b).abs().max() / b.abs().max()) <= 1e-13
    assert sum(counts) == sc["M"] and max(counts) - min(counts) <= 1


# ---------------------------------------------------------------------------------------------------------------------
# losses that mix the fused cost with the caller's own terms on states / inputs
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fused+states+inputs", "inputs-only-user-cost", "fused-unused-inputs-only"])
def test_mixed_losses_vs_oracle_autograd(nh, mode):
    """(cost + f(states) + g(inputs)).backward(): the fused cost's gradient and the caller's grad_states / grad_inputs enter the same
    backward sweep and must ADD (they used to be exclusive); a user cost that depends on the inputs alone must be accepted."""
    import scenarios
    from oracle import mcpilco_oracle as O
    sc = scenarios.scenario("c1")
    T = Hh.T
    pol_o = Hh.oracle_policy(sc, True)
    x0 = O.initial_particles(T(sc["x0_mean"]), T(sc["x0_var"]), T(sc["eps0"]))
    st, inp = O.rollout(Hh.oracle_model(sc), Hh.oracle_fit(sc), pol_o, x0, T(sc["eps"]), T(sc["masks"]), sc["p_dropout"])
    cost_o, _ = O.expected_cost(Hh.oracle_cost(sc, st))
    AB, obj, dev = _build_obj(sc)
    if mode == "inputs-only-user-cost":
        import mcpilco_b200.policy_learning.Cost_function as CF
        obj.cost_function = CF.Expected_cost(lambda states, inputs, trial_index: (inputs ** 2).sum(-1))
    G = AB.tensor_factory(dev)
    noise = dict(eps0=G(sc["eps0"]), eps=G(sc["eps"]), masks=G(sc["masks"]))
    states, inputs = obj.apply_policy(**AB.apply_kwargs(sc, dev), _noise=noise)
    cost, _ = obj.cost_function(states, inputs, 0)
    if mode == "fused+states+inputs":
        loss_o = cost_o + 0.3 * (st ** 2).mean() + 0.2 * (inp ** 2).mean()
        loss = cost + 0.3 * (states ** 2).mean() + 0.2 * (inputs ** 2).mean()
    elif mode == "inputs-only-user-cost":
        loss_o = (inp ** 2).sum(-1).mean(1).sum()
        loss = cost
    else:
        loss_o = 0.2 * (inp ** 2).mean()
        loss = 0.2 * (inputs ** 2).mean()
    loss_o.backward()
    loss.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-6 * abs(float(loss_o))
    pol = obj.control_policy
    for a, b in ((pol.log_lengthscales, pol_o["log_ls"]), (pol.centers, pol_o["centers"]), (pol.f_linear.weight, pol_o["W"])):
        assert relmax(a.grad, b.grad.numpy()) < REL_GRAD


# ---------------------------------------------------------------------------------------------------------------------
# the optimisation step with the rollout's forward + backward captured in a CUDA graph
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,numpy_init", [("c1", False), ("c3", True), ("c4", False)])
def test_graphed_optimisation_step_is_bit_identical_to_the_eager_loop(nh, monkeypatch, capsys, name, numpy_init):
    """reinforce_policy replays ONE captured graph per step (fresh Philox key through the device seed word) instead of enqueueing a
    few hundred launches: same keys, same kernels, same Adam updates, so the cost / std histories and the final trajectories must be
    bit-identical to the un-graphed loop (MCPILCO_NO_GRAPH=1).  numpy_init: initial distribution given as host arrays, like the
    reference's scripts do (they must be staged outside the captured region)."""
    import scenarios
    import mcpilco_b200.policy_learning.MC_PILCO as MCP
    sc = scenarios.scenario(name)
    outs = {}
    for mode in ("graph", "eager"):
        if mode == "eager":
            monkeypatch.setenv("MCPILCO_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("MCPILCO_NO_GRAPH", raising=False)
        AB, obj, dev = _build_obj(sc)
        T = AB.tensor_factory(dev)
        conv = (lambda a: np.asarray(a)) if numpy_init else T
        built = []
        orig = MCP._GraphedRollout.__init__
        monkeypatch.setattr(MCP._GraphedRollout, "__init__", lambda self, *a, **k: (built.append(1), orig(self, *a, **k))[1])
        torch.manual_seed(0)
        out = obj.reinforce_policy(T_control=sc["H"] * obj.T_sampling + 1e-9, num_particles=48, trial_index=0,
                                   particles_initial_state_mean=conv(sc["x0_mean"]), particles_initial_state_var=conv(sc["x0_var"]),
                                   flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None,
                                   flg_particles_init_multi_gauss=False, opt_steps_list=[12], lr_list=[0.05],
                                   f_optimizer="lambda p, lr : torch.optim.Adam(p, lr)", num_step_print=100, p_dropout_list=[0.1])
        monkeypatch.setattr(MCP._GraphedRollout, "__init__", orig)
        assert len(built) == (1 if mode == "graph" else 0)          # captured once, replayed 12 times
        outs[mode] = out
        assert np.isfinite(out[0]).all() and out[0].shape == (12,)
    capsys.readouterr()
    for a, b in zip(outs["graph"], outs["eager"]):
        assert a.shape == b.shape and np.array_equal(a, b)

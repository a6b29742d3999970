"""GPU parity: the CUDA path (through the C ABI) against the golden vectors produced by the reference and
against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): posterior mean/var and rollout cost 1e-5 relative, policy
gradients 1e-4 relative (fp64 mode).  Where the check is formula-level (the reference's own alpha/K^-1
are fed in) the tolerance is tightened to rounding level.
"""
import numpy as np
import pytest
import torch

import helpers as Hh
import scenarios
from oracle import mcpilco_oracle as O

pytestmark = pytest.mark.gpu

REL_VAL = 1e-5   # north-star tolerance for mean / var / cost
REL_GRAD = 1e-4  # north-star tolerance for policy gradients


@pytest.fixture(scope="module")
def nh():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import native_helpers
    return native_helpers


def close(a, b, rtol, atol=0.0):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), rtol=rtol, atol=atol)


def relmax(a, b):
    """max |a-b| / max |b|: the norm-wise relative error used for gradient tensors."""
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", scenarios.ALL)
def test_covariance(nh, name):
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X, Xs = nh.G(sc["X"]), nh.G(g["Xs"])
    for e, sp in enumerate(nh.native_specs(sc)):
        close(ops.gp_covariance(sp, Xs, X), g[f"Kss_{e}"], 1e-12, 1e-15)
        close(ops.gp_covariance(sp, X, None, add_noise=True), g[f"Knoise_{e}"], 1e-12, 1e-15)
        close(ops.gp_diag_covariance(sp, Xs), g[f"kdiag_{e}"], 1e-12)


@pytest.mark.parametrize("name", scenarios.ALL)
def test_precompute(nh, name):
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X = nh.G(sc["X"])
    for e, sp in enumerate(nh.native_specs(sc)):
        alpha, Kinv, Lf = ops.gp_precompute(sp, X, nh.G(sc["Y"][:, e:e + 1]), want_L=True)
        # conditioning-limited (cond(K) ~ 1e5..1e6): same bound the oracle is held to against the reference
        close(Kinv, g[f"Kinv_{e}"], 1e-7, 1e-7 * np.abs(g[f"Kinv_{e}"]).max())
        close(alpha, g[f"alpha_{e}"], 1e-7, 1e-7 * np.abs(g[f"alpha_{e}"]).max())
        Kn = torch.tensor(g[f"Knoise_{e}"], dtype=torch.float64, device=X.device)
        close(Lf @ Lf.t(), g[f"Knoise_{e}"], 1e-12, 1e-14)                      # L L^T = K
        close(Kinv @ Kn, np.eye(X.shape[0]), 0, 1e-8)                            # K^-1 K = I
        assert torch.equal(Kinv, Kinv.t())                                      # exactly symmetric


@pytest.mark.parametrize("name", scenarios.ALL)
def test_posterior_formula(nh, name):
    """Reference alpha / K^-1 in, posterior out: formula parity at rounding level."""
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    gps = nh.native_fit(sc, golden=g)
    mean, var = ops.gp_predict(gps, nh.G(g["Xs"]))
    for e in range(sc["E"]):
        close(mean[:, e:e + 1], g[f"pmean_{e}"], 1e-10, 1e-13)
        close(var[:, e], g[f"pvar_{e}"], 1e-8, 1e-13)


@pytest.mark.parametrize("name", scenarios.ALL)
def test_posterior_end_to_end(nh, name):
    """Own precompute + own posterior vs the reference, at the stated tolerance."""
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    mean, var = ops.gp_predict(nh.native_fit(sc), nh.G(g["Xs"]))
    for e in range(sc["E"]):
        close(mean[:, e:e + 1], g[f"pmean_{e}"], REL_VAL, 1e-9)
        close(var[:, e], g[f"pvar_{e}"], REL_VAL)


@pytest.mark.parametrize("name", ["c1", "c4", "delta"])
def test_posterior_jacobian(nh, name):
    """d mean / d x and d var / d x against autograd through the oracle."""
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    gps = nh.native_fit(sc, golden=g)
    Xs = nh.G(g["Xs"])
    _, _, jm, jv = ops.gp_predict(gps, Xs, jac=True)
    X = Hh.T(sc["X"])
    for e, sp in enumerate(Hh.oracle_specs(sc)):
        xs = Hh.T(g["Xs"]).requires_grad_(True)
        mu, var = O.gp_predict(sp, X, Hh.T(g[f"alpha_{e}"]), Hh.T(g[f"Kinv_{e}"]), xs)
        gm, = torch.autograd.grad(mu.sum(), xs, retain_graph=True)
        gv, = torch.autograd.grad(var.sum(), xs)
        assert relmax(jm[:, e, :], gm.numpy()) < 1e-9
        assert relmax(jv[:, e, :], gv.numpy()) < 1e-7


def _run_rollout(nh, sc, gps, fused_cost=True):
    plan, ptens = nh.native_plan(sc, gps, need_grad=True, fused_cost=fused_cost)
    states, inputs = plan.forward(nh.x0_of(sc))
    return plan, states, inputs


@pytest.fixture(params=["persistent", "fused-small", "per-step"])
def rollout_path(request, monkeypatch):
    """Cart-pole-sized rollouts (D <= 6) take the whole-horizon persistent cluster kernel, other small rollouts (and these with
    MCPILCO_NO_PERSIST=1) the fused two-launch-per-step kernels; MCPILCO_NO_SMALL_PATH=1 sends the same rollout through the per-step
    kernels the large shapes use.  All are held to the same golden vectors."""
    for v in ("MCPILCO_NO_SMALL_PATH", "MCPILCO_NO_PERSIST", "MCPILCO_PERSIST"):
        monkeypatch.delenv(v, raising=False)
    if request.param == "persistent":
        monkeypatch.setenv("MCPILCO_PERSIST", "1")   # wherever eligible, not only where the default heuristic picks it
    elif request.param == "per-step":
        monkeypatch.setenv("MCPILCO_NO_SMALL_PATH", "1")
    elif request.param == "fused-small":
        monkeypatch.setenv("MCPILCO_NO_PERSIST", "1")
    return request.param


@pytest.mark.parametrize("name", scenarios.ALL)
def test_rollout_formula(nh, name, rollout_path):
    """Reference factors in; trajectories, cost and gradients out."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    plan, states, inputs = _run_rollout(nh, sc, nh.native_fit(sc, golden=g))
    close(states, g["states"], 1e-8, 1e-11)
    close(inputs, g["inputs"], 1e-8, 1e-11)
    close(plan.cost_out[0], g["cost"], 1e-10)
    close(plan.cost_out[1], g["std_cost"], 1e-9)
    gr = plan.backward(grad_cost=1.0)
    assert relmax(gr["log_ls"], g["g_log_ls"]) < 1e-7
    assert relmax(gr["centers"], g["g_centers"]) < 1e-7
    assert relmax(gr["W"], g["g_W"]) < 1e-7
    if "g_bias" in g:
        assert relmax(gr["bias"], g["g_bias"]) < 1e-7


@pytest.mark.parametrize("name", scenarios.ALL)
def test_rollout_end_to_end(nh, name, rollout_path):
    """Own precompute; north-star tolerances."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    plan, states, inputs = _run_rollout(nh, sc, nh.native_fit(sc))
    close(plan.cost_out[0], g["cost"], REL_VAL)
    close(plan.cost_out[1], g["std_cost"], 1e-4)
    assert relmax(states, g["states"]) < REL_VAL
    gr = plan.backward(grad_cost=1.0)
    for k in ("log_ls", "centers", "W"):
        assert relmax(gr[k], g["g_" + k]) < REL_GRAD


@pytest.mark.parametrize("name", ["c1", "c3", "c4"])
def test_rollout_generic_cost_path(nh, name):
    """No fused cost: the caller differentiates its own cost w.r.t. states and hands grad_states back."""
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    plan, states, inputs = _run_rollout(nh, sc, nh.native_fit(sc, golden=g), fused_cost=False)
    st = states.detach().cpu().requires_grad_(True)
    cost, _ = O.expected_cost(Hh.oracle_cost(sc, st))
    cost.backward()
    gr = plan.backward(grad_states=st.grad.to(states.device))
    for k in ("log_ls", "centers", "W"):
        assert relmax(gr[k], g["g_" + k]) < 1e-7


def test_backward_finite_differences(nh):
    """Adjoint kernel in isolation: central differences of the CUDA forward cost (T6 of SURVEY.md §4)."""
    sc, g = scenarios.scenario("c1"), Hh.load_golden("c1")
    gps = nh.native_fit(sc, golden=g)
    plan, ptens = nh.native_plan(sc, gps, need_grad=True)
    x0 = nh.x0_of(sc)
    plan.forward(x0)
    gr = plan.backward(grad_cost=1.0, want_gx0=True)
    rs = np.random.RandomState(0)
    for key, gk in (("centers", "centers"), ("W", "W"), ("log_ls", "log_ls")):
        t = ptens[key]
        flat = t.view(-1)
        for i in rs.choice(flat.numel(), size=min(4, flat.numel()), replace=False):
            h = 1e-6
            old = float(flat[i])
            flat[i] = old + h; plan.forward(x0); cp = float(plan.cost_out[0])
            flat[i] = old - h; plan.forward(x0); cm = float(plan.cost_out[0])
            flat[i] = old
            fd = (cp - cm) / (2 * h)
            an = float(gr[gk].view(-1)[i])
            assert abs(fd - an) <= 2e-5 * max(abs(an), 1e-3), (key, int(i), fd, an)
    # gradient w.r.t. the initial particles
    i, j, h = 3, 1, 1e-6
    xp = x0.clone(); xp[i, j] += h; plan.forward(xp); cp = float(plan.cost_out[0])
    xm = x0.clone(); xm[i, j] -= h; plan.forward(xm); cm = float(plan.cost_out[0])
    an = float(gr["x0"][i, j])
    assert abs((cp - cm) / (2 * h) - an) <= 2e-5 * max(abs(an), 1e-3)


def test_philox_rollout_statistics(nh):
    """Production noise (counter-based Philox): deterministic per seed, different across seeds, dropout rate and
    N(0,1) moments as specified; sharding-invariant by construction (keyed by global particle id)."""
    sc, g = scenarios.scenario("c2"), Hh.load_golden("c2")
    sc = dict(sc); sc["M"] = 4096
    gps = nh.native_fit(sc, golden=g)
    x0 = nh.G(sc["x0_mean"]).repeat(sc["M"], 1)
    outs = []
    for seed in (1, 1, 2):
        plan, _ = nh.native_plan(sc, gps, need_grad=False, inject=False, seed=seed)
        s, u = plan.forward(x0)
        outs.append((s.clone(), u.clone(), float(plan.cost_out[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert not torch.equal(outs[0][0], outs[2][0])
    # two half-size shards with particle offsets reproduce the full run bit for bit
    half = sc["M"] // 2
    sh = dict(sc); sh["M"] = half
    parts = []
    for r in range(2):
        plan, _ = nh.native_plan(sh, gps, need_grad=False, inject=False, seed=1, particle_offset=r * half, M_global=sc["M"])
        s, _ = plan.forward(x0[r * half:(r + 1) * half])
        parts.append(s.clone())
    assert torch.equal(torch.cat(parts, 1), outs[0][0])


@pytest.mark.parametrize("Nn", [2048, 8192])
def test_size_independent_properties_large(nh, Nn):
    """Sweep-sized GPs (N = 2048 and the bench's full N = 8192): identities that hold for any size, checked on the CUDA results alone.
    At a training input x_i:  mean = y_i - sn2 * alpha_i  and  var = sn2 * (1 - sn2 * Kinv_ii)."""
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import _pack as P
    gen = torch.Generator().manual_seed(0)
    X, Y = O.cartpole_dataset(Nn, 0.1, gen)
    spec = P.spec_from_dict({"D": 6, "log_ls": [2, 2, 2, 0.8, 1.5, 2.5], "lambda": 1.0, "mean": 0.0,
                             "mpk": [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)], "sigma_n": 0.1})
    Xg, yg = X.to("cuda:0"), Y[:, 0:1].contiguous().to("cuda:0")
    alpha, Kinv = ops.gp_precompute(spec, Xg, yg)
    K = ops.gp_covariance(spec, Xg, None, add_noise=True)
    resid = (K @ alpha - yg).abs().max().item()
    ident = (Kinv @ K - torch.eye(Nn, dtype=torch.float64, device="cuda:0")).abs().max().item()
    print("N=%d  |K alpha - y| = %.3e  |Kinv K - I| = %.3e" % (Nn, resid, ident))
    assert resid < 1e-10 and ident < 1e-10, (resid, ident)      # measured: 8.5e-12 / 1.7e-12 at N = 8192
    gp = ops.FittedGp(spec, Xg, alpha, Kinv)
    mean, var = ops.gp_predict([gp], Xg[:512])      # 512 x N: the TMA-pipelined GEMM at N = 8192
    sn2 = 0.01
    close(mean[:, 0], (yg[:512, 0] - sn2 * alpha[:512, 0]).cpu().numpy(), 1e-7, 1e-9)
    close(var[:, 0], (sn2 * (1 - sn2 * torch.diagonal(Kinv)[:512])).cpu().numpy(), 1e-5, 1e-10)
    # and against the oracle formula on a handful of fresh points (CPU finishes in seconds at this size)
    Xs = X[:16] + 0.05
    sp_o = O.make_spec(6, log_ls=[2, 2, 2, 0.8, 1.5, 2.5], mpk_log_pars=[[-5, -5, -5, -4, -4, -4, -3.0], [-5, -5, -4, -2, -1, -4.0] * 2],
                       sigma_n=0.1)
    mo, vo = O.gp_predict(sp_o, X, alpha.cpu(), Kinv.cpu().contiguous(), Xs)
    mg, vg = ops.gp_predict([gp], Xs.to("cuda:0"))
    close(mg, mo.numpy(), 1e-9, 1e-12)
    close(vg[:, 0], vo.numpy(), REL_VAL)


def test_headline_shape_rollout_vs_oracle(nh, monkeypatch):
    """The bench's kernels (per-step path: cov_fast, TMA GEMM, fast reduce, rollout_bwd) at the bench's FULL training size N = 8192,
    E = 2, SE + MPK(2), against the CPU oracle on the same factors: trajectories, cost and policy gradients at north-star tolerances.
    Few particles and a short horizon keep the oracle (2 x 3.4e10-flop GEMMs per output and step, plus autograd) to seconds."""
    monkeypatch.setenv("MCPILCO_NO_SMALL_PATH", "1")
    sc = scenarios.headline(N=8192, M=256, H=3)
    gps = nh.native_fit(sc)
    ogps = [(sp, Hh.T(sc["X"]), g.alpha.cpu().reshape(-1, 1), g.Kinv.cpu().contiguous()) for sp, g in zip(Hh.oracle_specs(sc), gps)]
    ref = Hh.oracle_rollout(sc, gps=ogps)
    plan, ptens = nh.native_plan(sc, gps, need_grad=True)
    states, inputs = plan.forward(nh.x0_of(sc))
    assert relmax(states, ref["states"]) < REL_VAL and relmax(inputs, ref["inputs"]) < REL_VAL
    close(plan.cost_out[0], ref["cost"], REL_VAL)
    close(plan.cost_out[1], ref["std_cost"], 1e-4)
    gr = plan.backward(grad_cost=1.0)
    for k in ("log_ls", "centers", "W"):
        assert relmax(gr[k], ref["g_" + k]) < REL_GRAD, k


def test_edge_cases(nh):
    from mcpilco_b200 import _native as Nn
    from mcpilco_b200 import _ops as ops
    sc, g = scenarios.scenario("c2"), Hh.load_golden("c2")
    gps = nh.native_fit(sc, golden=g)
    # empty test set
    mean, var = ops.gp_predict(gps, torch.empty(0, sc["D"], dtype=torch.float64, device="cuda:0"))
    assert mean.shape == (0, 2) and var.shape == (0, 2)
    # single particle, horizon 1 (only the initial policy evaluation)
    s1 = dict(sc); s1["M"], s1["H"] = 1, 1
    s1["eps"], s1["masks"] = sc["eps"][:0, :1], sc["masks"][:1, :1]
    plan, _ = nh.native_plan(s1, gps, need_grad=False)
    st, inp = plan.forward(nh.x0_of(sc)[:1])
    close(st[0], g["states"][0, :1], 1e-12)
    close(inp[0], g["inputs"][0, :1], 1e-9)
    # N = 1 training point, odd N
    X1 = nh.G(sc["X"][:1]); y1 = nh.G(sc["Y"][:1, :1])
    sp = nh.native_specs(sc)[0]
    a1, K1 = ops.gp_precompute(sp, X1, y1)
    k = 1.0 + np.exp(-4.2) ** 2
    close(K1, [[1.0 / k]], 1e-13); close(a1, sc["Y"][:1, :1] / k, 1e-13)
    X5 = nh.G(sc["X"][:5]); y5 = nh.G(sc["Y"][:5, :1])
    a5, K5 = ops.gp_precompute(sp, X5, y5)
    Kn = ops.gp_covariance(sp, X5, None, add_noise=True)
    close(K5 @ Kn, np.eye(5), 0, 1e-10)
    # errors are loud: CPU tensors are rejected, bad shapes return MCP_E_ARG
    with pytest.raises(RuntimeError):
        ops.gp_covariance(sp, torch.zeros(3, 6, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        ops.gp_predict(gps, torch.zeros(3, 5, dtype=torch.float64, device="cuda:0"))
    # non-SPD input propagates NaN instead of raising (MC_PILCO.py:451,497 rely on it)
    bad = P_bad_spec(nh)
    a, K = ops.gp_precompute(bad, X5, y5)
    assert torch.isnan(a).any() or torch.isinf(a).any() or (a.abs() > 1e6).any()


def P_bad_spec(nh):
    from mcpilco_b200 import _pack as P
    s = P.spec_from_dict({"D": 6, "log_ls": [50.0] * 6, "lambda": 1.0, "mean": 0.0, "mpk": [], "sigma_n": 0.0})
    return s  # lengthscales so long that K is numerically rank one and sigma_n = 0: Cholesky breaks down


@pytest.mark.parametrize("N,M", [(300, 77), (257, 130), (1000, 513), (4096, 300), (2003, 800)])
def test_posterior_tiles_and_tails(nh, N, M):
    """TMA-pipelined FP64 tensor-core contraction at ragged sizes (row / column / K tails are zero-filled by the TMA unit):
    posterior from the CUDA path vs the same formula evaluated with torch fp64 matmul on the CUDA covariance matrices."""
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import _pack as P
    rs = np.random.RandomState(N + M)
    spec = P.spec_from_dict({"D": 6, "log_ls": [2, 2, 2, 0.8, 1.5, 2.5], "lambda": 1.0, "mean": 0.02,
                             "mpk": [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)], "sigma_n": 0.1})
    X = nh.G(rs.uniform(-2, 2, (N, 6)))
    y = nh.G(rs.randn(N, 1))
    Xs = nh.G(rs.uniform(-2, 2, (M, 6)))
    alpha, Kinv = ops.gp_precompute(spec, X, y)
    mean, var, jm, jv = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv)], Xs, jac=True)
    Ks = ops.gp_covariance(spec, Xs, X)
    V = Ks @ Kinv
    mean_ref = 0.02 + Ks @ alpha
    q_ref = (V * Ks).sum(1)
    var_ref = ops.gp_diag_covariance(spec, Xs) - q_ref
    close(mean, mean_ref.cpu().numpy(), 1e-10, 1e-10)  # |alpha| ~ 1e2..1e3 over N terms: absolute rounding floor
    # the quadratic form itself agrees to rounding; var = k** - q inherits the cancellation (var/k** ~ 1e-2..1e-4 here)
    q = ops.gp_diag_covariance(spec, Xs) - var[:, 0]
    close(q, q_ref.cpu().numpy(), 1e-10)  # N-term sums in different orders (DMMA tiles vs cuBLAS)
    # var = k** - q: its error is q's (1e-10 relative to q, not to the much smaller var)
    assert float(((var[:, 0] - var_ref).abs() - (1e-10 * q_ref.abs() + 1e-12)).max()) <= 0.0
    assert torch.isfinite(jm).all() and torch.isfinite(jv).all()


@pytest.mark.parametrize("N,M", [(1, 3), (61, 40), (300, 77), (1000, 513), (2003, 800), (4096, 300)])
def test_forward_only_triangular_form_matches_full_product(nh, N, M):
    """No Jacobians wanted (no-grad rollouts, reference MC_PILCO.py:430-456; get_estimate_from_alpha): with the triangular factor
    R = L^-1 from the precompute the posterior is  var = k** - |R k*|^2  through a contraction trimmed to the triangle (half the
    flops); it must agree with the reference's full form  k** - k*^T Kinv k*  (and R must be the inverse factor of Kinv)."""
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import _pack as P
    rs = np.random.RandomState(N + M)
    spec = P.spec_from_dict({"D": 6, "log_ls": [2, 2, 2, 0.8, 1.5, 2.5], "lambda": 1.0, "mean": 0.02,
                             "mpk": [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)], "sigma_n": 0.1})
    X = nh.G(rs.uniform(-2, 2, (N, 6)))
    y = nh.G(rs.randn(N, 1))
    Xs = nh.G(rs.uniform(-2, 2, (M, 6)))
    alpha, Kinv, R = ops.gp_precompute(spec, X, y, want_Linv=True)
    assert torch.equal(R, torch.tril(R))
    close(R.t() @ R, Kinv.cpu().numpy(), 1e-9, 1e-9 * float(Kinv.abs().max()))
    m_full, v_full = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv)], Xs)
    m_tri, v_tri = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv, Linv=R)], Xs)
    close(m_tri, m_full.cpu().numpy(), 1e-10, 1e-10)
    kd = ops.gp_diag_covariance(spec, Xs)
    close(kd - v_tri[:, 0], (kd - v_full[:, 0]).cpu().numpy(), 1e-9)   # the quadratic forms agree to rounding
    close(v_tri, v_full.cpu().numpy(), 1e-6, 1e-12)
    assert (v_tri > 0).all()
    # with Jacobians the full product is taken either way: bit-identical
    a = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv, Linv=R)], Xs, jac=True)
    b = ops.gp_predict([ops.FittedGp(spec, X, alpha, Kinv)], Xs, jac=True)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    # the class API hands the factor along on the K_X_inv tensor it returns
    assert ops.attached_linv(Kinv) is None
    ops.attach_linv(Kinv, R)
    assert ops.attached_linv(Kinv) is R and ops.FittedGp(spec, X, alpha, Kinv).Linv is not None
    Kinv.mul_(1.0)                                                    # modified in place: the attachment is no longer trusted
    assert ops.attached_linv(Kinv) is None


def test_cost_stats_output_and_shard_merge(nh):
    """The per-step {mean, M2} statistics the rollout exports are what two half-size shards need to rebuild the global
    Expected_cost (single GPU here: the two shards run one after the other; the collectives are covered by the gloo test)."""
    from mcpilco_b200 import distributed as D
    sc, g = scenarios.scenario("c2"), Hh.load_golden("c2")
    sc = dict(sc); sc["M"] = 1000
    gps = nh.native_fit(sc, golden=g)
    x0 = nh.G(sc["x0_mean"]).repeat(sc["M"], 1)
    full, _ = nh.native_plan(sc, gps, need_grad=False, inject=False, seed=5)
    full.forward(x0)
    costs = full.costs
    close(full.cost_stats[:, 0], costs.mean(1).cpu().numpy(), 1e-13)
    close(full.cost_stats[:, 1], ((costs - costs.mean(1, keepdim=True)) ** 2).sum(1).cpu().numpy(), 1e-10, 1e-20)
    stats, counts = [], []
    for r in range(3):
        off, cnt = D.shard(sc["M"], r, 3)
        sh = dict(sc); sh["M"] = cnt
        plan, _ = nh.native_plan(sh, gps, need_grad=False, inject=False, seed=5, particle_offset=off, M_global=sc["M"])
        plan.forward(x0[off:off + cnt])
        stats.append(plan.cost_stats.clone()); counts.append(cnt)
    mean, m2 = D.merge_cost_stats(torch.stack(stats), counts)
    cost, std = D.expected_cost_from_stats(mean, m2, sc["M"])
    close(cost, float(full.cost_out[0]), 1e-13)
    close(std, float(full.cost_out[1]), 1e-10)


@pytest.mark.parametrize("N", [40, 96, 200, 330])
def test_ur5_true_dimensions_vs_oracle(nh, N):
    """Config 4 at its real dimensions (D = 24, E = 6, Ds = 12, Du = 6): trajectories, cost and policy gradients against the CPU oracle
    (own precompute on both sides), plus the single-step posterior and its Jacobians.  N = 40: one 64-point tile, no cluster split in the
    wide reduce; 96: two tiles, clusters of two; 200: four tiles with a ragged tail, clusters of four; 330: six tiles, two of the four CTAs
    take a second one."""
    from mcpilco_b200 import _ops as ops
    sc = scenarios.ur5_full(N=N)
    ref = Hh.oracle_rollout(sc)
    gps = nh.native_fit(sc)
    plan, _ = nh.native_plan(sc, gps, need_grad=True)
    states, inputs = plan.forward(nh.x0_of(sc))
    assert relmax(states, ref["states"]) < REL_VAL and relmax(inputs, ref["inputs"]) < REL_VAL
    close(plan.cost_out[0], ref["cost"], REL_VAL)
    gr = plan.backward(grad_cost=1.0)
    for k in ("log_ls", "centers", "W"):
        assert relmax(gr[k], ref["g_" + k]) < REL_GRAD, k
    # posterior Jacobians of every output at D = 24 against oracle autograd
    ogps = Hh.oracle_fit(sc)
    Xs = sc["X"][:7] + 0.05
    _, _, jm, jv = ops.gp_predict(gps, nh.G(Xs), jac=True)
    for e, (sp, X, alpha, Kinv) in enumerate(ogps):
        xs = Hh.T(Xs).requires_grad_(True)
        mu, var = O.gp_predict(sp, X, alpha, Kinv, xs)
        gm, = torch.autograd.grad(mu.sum(), xs, retain_graph=True)
        gv, = torch.autograd.grad(var.sum(), xs)
        assert relmax(jm[:, e, :], gm.numpy()) < 1e-6 and relmax(jv[:, e, :], gv.numpy()) < 1e-5


@pytest.mark.parametrize("D,linear", [(9, True), (13, False), (16, True), (24, False), (31, True), (32, True)])
def test_wide_posterior_other_widths_vs_oracle(nh, D, linear):
    """The wide-input posterior (K* kernel, contraction, DMMA reduce over clusters) at input widths other than the UR5's 24 — not a
    multiple of 4 or 8, the maximum 32 — with and without the linear term, particle counts that leave the last CTA (8 particles)
    and the last warp (2) ragged: mean, variance and both Jacobians against the oracle and its autograd."""
    from mcpilco_b200 import _ops as ops, _pack as P
    rs = np.random.RandomState(100 + D)
    N, M = 150, 21
    X = rs.uniform(-1.5, 1.5, (N, D))
    w = rs.randn(D) / np.sqrt(D)
    Y = (np.sin(X @ w) + 0.3 * (X @ rs.randn(D)) / np.sqrt(D) + 0.01 * rs.randn(N))[:, None]
    g = {"D": D, "log_ls": np.log(2.5) + 0.1 * rs.randn(D), "lambda": 1.3, "sigma_n": 0.05, "mean": 0.1,
         "mpk": [0.1 * np.exp(0.1 * rs.randn(D + 1))] if linear else []}
    sp = P.spec_from_dict(g)
    alpha, Kinv = ops.gp_precompute(sp, nh.G(X), nh.G(Y))
    gp = ops.FittedGp(sp, nh.G(X), alpha, Kinv)
    Xs = rs.uniform(-1.5, 1.5, (M, D))
    mu, var, jm, jv = ops.gp_predict([gp], nh.G(Xs), jac=True)
    # oracle on the same factors
    sc = {"D": D, "X": X, "Y": Y, "gps": [g]}
    osp2, oX, oalpha, oKinv = Hh.oracle_fit(sc)[0]
    xs = Hh.T(Xs).requires_grad_(True)
    omu, ovar = O.gp_predict(osp2, oX, oalpha, oKinv, xs)
    gm, = torch.autograd.grad(omu.sum(), xs, retain_graph=True)
    gv, = torch.autograd.grad(ovar.sum(), xs)
    assert relmax(mu[:, 0], omu.detach().numpy()) < 1e-8 and relmax(var[:, 0], ovar.detach().numpy()) < 1e-5
    assert relmax(jm[:, 0, :], gm.numpy()) < 1e-6 and relmax(jv[:, 0, :], gv.numpy()) < 1e-5


def test_wide_covariance_kernel_is_bit_identical_to_generic(nh):
    """Wide gp inputs (D = 24) take a dedicated K* kernel; the K + noise build of the precompute takes the generic one.  Same arithmetic
    per entry: off the diagonal the two must agree bit for bit (and on it up to the added noise)."""
    from mcpilco_b200 import _ops as ops
    sc = scenarios.ur5_full()
    X = nh.G(sc["X"])
    for sp in nh.native_specs(sc):
        Kw = ops.gp_covariance(sp, X, X)                       # cross-covariance path -> cov_wide_kernel
        Kg = ops.gp_covariance(sp, X, None, add_noise=True)    # K + sigma_n^2 I    -> cov_kernel
        off = ~torch.eye(X.shape[0], dtype=torch.bool, device=X.device)
        assert torch.equal(Kw[off], Kg[off])
        assert torch.allclose(torch.diagonal(Kg) - torch.diagonal(Kw), torch.full((X.shape[0],), sp.sigma_n2, dtype=torch.float64, device=X.device),
                              rtol=0, atol=1e-15)
        # ragged shapes: rows / columns that are not multiples of the strip sizes
        Xa, Xb = X[:37], X[5:96]
        Kr = ops.gp_covariance(sp, Xa, Xb)
        assert torch.equal(Kr, Kw[:37, 5:96])


def test_ur5_batched_step_is_bit_identical_to_per_output_chains(nh, monkeypatch):
    """Wide-input rollouts batch all outputs of a step into three launches (MCPILCO_NO_BATCHED_STEP=1 restores the per-output chains
    on side streams); same device code, so trajectories, cost and gradients must agree bit for bit."""
    sc = scenarios.ur5_full()
    gps = nh.native_fit(sc)
    out = {}
    for mode in ("batched", "chains"):
        if mode == "chains":
            monkeypatch.setenv("MCPILCO_NO_BATCHED_STEP", "1")
        else:
            monkeypatch.delenv("MCPILCO_NO_BATCHED_STEP", raising=False)
        plan, _ = nh.native_plan(sc, gps, need_grad=True)
        states, inputs = plan.forward(nh.x0_of(sc))
        gr = plan.backward(grad_cost=1.0)
        out[mode] = (states.clone(), inputs.clone(), plan.cost_out.clone(), {k: v.clone() for k, v in gr.items() if v is not None})
    a, b = out["batched"], out["chains"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k


# ---------------------------------------------------------------------------------------------------------------------
# opt-in INT8 tensor-core contraction with error compensation.  Stated tolerances: with 8 digit planes the
# posterior variance agrees with the native fp64 path to 1e-7 relative, with 7 planes to 1e-5 (the north-star bound); the
# contraction itself is exact up to the 2^-63 / 2^-55 representation of the operands relative to their row maxima.
# ---------------------------------------------------------------------------------------------------------------------
def _need_ozaki():
    from mcpilco_b200 import _native as Nn
    if not Nn.lib().mcpilco_ozaki_available():
        pytest.skip("built without CUTLASS headers")


@pytest.mark.parametrize("M,N", [(128, 128), (77, 300), (513, 1000), (64, 8320)])  # the last one needs two K segments
def test_ozaki_contraction_matches_fp64_gemm(nh, M, N):
    from mcpilco_b200 import _ops as ops
    _need_ozaki()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g) * torch.logspace(-3, 3, M, dtype=torch.float64, device="cuda")[:, None]
    B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g) * torch.logspace(2, -2, N, dtype=torch.float64, device="cuda")[:, None]
    ref = A @ B.t()
    scale = A.abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).t() * N   # the error bound is relative to rowmax * colmax * K
    # 8 planes: indistinguishable from the fp64 reference's own rounding (~2^-53 of the scale); 7 planes: 2^-55 operand truncation
    for S, tol in ((8, 2.0 ** -52), (7, 2.0 ** -49)):
        V, _, _ = ops.ozaki_matmul(A, B, S)
        assert float(((V - ref).abs() / scale).max()) < tol, S


@pytest.mark.parametrize("N,M", [(1000, 513), (2048, 700)])
def test_ozaki_posterior_within_stated_tolerance(nh, N, M):
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import _pack as P
    _need_ozaki()
    rs = np.random.RandomState(N)
    spec = P.spec_from_dict({"D": 6, "log_ls": [2, 2, 2, 0.8, 1.5, 2.5], "lambda": 1.0, "mean": 0.0,
                             "mpk": [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)], "sigma_n": 0.1})
    gen = torch.Generator().manual_seed(N)
    X, Y = O.cartpole_dataset(N, 0.1, gen)
    Xg, yg = X.to("cuda:0"), Y[:, 0:1].contiguous().to("cuda:0")
    alpha, Kinv = ops.gp_precompute(spec, Xg, yg)
    Xs = (X[torch.randint(0, N, (M,), generator=gen)] + 0.05 * torch.randn(M, 6, dtype=torch.float64, generator=gen)).to("cuda:0")
    m0, v0, jm0, jv0 = ops.gp_predict([ops.FittedGp(spec, Xg, alpha, Kinv, ozaki_slices=0)], Xs, jac=True)
    for S, tol in ((8, 1e-7), (7, 1e-5)):
        m1, v1, jm1, jv1 = ops.gp_predict([ops.FittedGp(spec, Xg, alpha, Kinv, ozaki_slices=S)], Xs, jac=True)
        assert torch.equal(m1, m0) and torch.equal(jm1, jm0)          # the mean does not go through the contraction
        assert float(((v1 - v0).abs() / v0.abs()).max()) < tol, (S, float(((v1 - v0).abs() / v0.abs()).max()))
        assert relmax(jv1, jv0.cpu().numpy()) < tol
    with pytest.raises(RuntimeError):
        ops.FittedGp(spec, Xg, alpha, Kinv, ozaki_slices=5)


@pytest.mark.parametrize("name", ["c1", "c3", "c4"])
def test_ozaki_rollout_against_goldens(nh, name, monkeypatch):
    """The whole rollout (per-step kernels, INT8 contraction with 8 planes) against the reference's trajectories, cost and gradients."""
    _need_ozaki()
    monkeypatch.setenv("MCPILCO_OZAKI", "8")
    monkeypatch.setenv("MCPILCO_NO_SMALL_PATH", "1")
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    gps = nh.native_fit(sc)
    assert all(gp.ozaki == 8 for gp in gps)
    plan, states, inputs = _run_rollout(nh, sc, gps)
    close(plan.cost_out[0], g["cost"], REL_VAL)
    assert relmax(states, g["states"]) < REL_VAL
    gr = plan.backward(grad_cost=1.0)
    for k in ("log_ls", "centers", "W"):
        assert relmax(gr[k], g["g_" + k]) < REL_GRAD


@pytest.mark.parametrize("kernel", ["se", "se+mpk"])
def test_sod_selection_vs_oracle(nh, kernel):
    """Device-side incremental greedy selection against the oracle's refit-per-candidate loop (reference GP_prior.py:232-257), in the
    natural order and in a given permutation."""
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import _pack as P
    rs = np.random.RandomState(11)
    N = 140
    sc = scenarios.scenario("c1")
    X = np.concatenate([sc["X"], sc["X"][rs.choice(48, N - 48)] + 0.3 * rs.randn(N - 48, 6)], 0)
    mpk = [np.exp([-5, -5, -5, -4, -4, -4, -3.0]), np.exp([-5, -5, -4, -2, -1, -4.0] * 2)] if kernel == "se+mpk" else []
    log_ls = [2, 2, 2, 0.8, 1.5, 2.5]
    spec = P.spec_from_dict({"D": 6, "log_ls": log_ls, "lambda": 1.0, "mean": 0.0, "mpk": mpk, "sigma_n": 0.05})
    so = O.make_spec(6, log_ls=log_ls, mpk_log_pars=[np.log(w) for w in mpk], sigma_n=0.05)
    thr = 0.1  # selects ~100 of the 140 points: many accept/reject decisions, some close to the threshold
    Y = Hh.T(rs.randn(N, 1))
    ref = O.sod_select(so, Hh.T(X), Y, thr)
    got = ops.gp_sod_select(spec, nh.G(X), thr)
    assert got == ref and 1 < len(ref) < N
    perm = [0] + (1 + rs.permutation(N - 1)).tolist()
    ref_p = [perm[i] for i in O.sod_select(so, Hh.T(X[perm]), Y, thr)]
    assert ops.gp_sod_select(spec, nh.G(X), thr, order=perm) == ref_p


@pytest.mark.parametrize("name", ["c1", "c4"])
def test_torch_custom_ops_match_goldens(nh, name):
    """torch.ops.mcpilco.* (torch.library layer over the C ABI) against the reference's golden vectors and the flat operators."""
    from mcpilco_b200 import _ops as ops
    from mcpilco_b200 import torch_ops as TO
    sc, g = scenarios.scenario(name), Hh.load_golden(name)
    X, Xs = nh.G(sc["X"]), nh.G(g["Xs"])
    for e, sp in enumerate(nh.native_specs(sc)):
        t, y = TO.spec_tensor(sp), nh.G(sc["Y"][:, e:e + 1])
        close(torch.ops.mcpilco.gp_covariance(t, Xs, X, False), g[f"Kss_{e}"], 1e-12, 1e-15)
        close(torch.ops.mcpilco.gp_covariance(t, X, None, True), g[f"Knoise_{e}"], 1e-12, 1e-15)
        close(torch.ops.mcpilco.gp_diag_covariance(t, Xs), g[f"kdiag_{e}"], 1e-12)
        alpha, Kinv = torch.ops.mcpilco.gp_precompute(t, X, y)
        close(alpha, g[f"alpha_{e}"], 1e-7, 1e-7 * np.abs(g[f"alpha_{e}"]).max())
        ga, gK = nh.G(g[f"alpha_{e}"]), nh.G(g[f"Kinv_{e}"])
        m, v, jm, jv = torch.ops.mcpilco.gp_predict_jac(t, X, ga, gK, Xs, 1.0)
        m2, v2 = torch.ops.mcpilco.gp_predict(t, X, ga, gK, Xs, 1.0)
        close(m, g[f"pmean_{e}"], 1e-10, 1e-13)
        close(v[:, 0], g[f"pvar_{e}"], 1e-8, 1e-13)
        rm, rv, rjm, rjv = ops.gp_predict([ops.FittedGp(sp, X, ga, gK)], Xs, jac=True)
        rm2, rv2 = ops.gp_predict([ops.FittedGp(sp, X, ga, gK)], Xs)
        assert torch.equal(m, rm) and torch.equal(v, rv) and torch.equal(m2, rm2) and torch.equal(v2, rv2)
        assert torch.equal(jm, rjm[:, 0, :]) and torch.equal(jv, rjv[:, 0, :])
        assert torch.equal(torch.ops.mcpilco.gp_nlml(t, X, y), ops.gp_nlml(sp, X, y))

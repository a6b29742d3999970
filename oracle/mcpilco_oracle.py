"""CPU oracle for the MC-PILCO particle-rollout hot path.

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch, functional restatement (torch, CPU,
float64) of the algorithm implemented by the reference's classes.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``; the product path (``mc-pilco_b200/``) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests/golden vectors of its own, so
the oracle is pinned against outputs of the reference itself: ``tests/golden/make_golden.py``
imports ``/root/reference`` (in the build container), runs its classes on seeded inputs with
injected noise and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

All ``file:line`` citations are relative to ``/root/reference``.  Operation order deliberately
follows the reference (matmul-form distances, explicit inverse, autograd graph kept alive) so that
timing this oracle on host cores is a fair stand-in for timing the reference ("kind": "port").

Data model (plain dicts, log-parameters exactly as the reference stores them):

gp spec  = {"D": int,
            "se":  None | {"active": LongTensor, "log_ls": [d], "log_lambda": [1], "mean": [1]},
            "mpk": [ {"active": LongTensor, "deg": int, "offset": bool, "log_par": [deg*(d+offset)]}, ...],
            "sigma_n_log": scalar tensor, "sigma_n_num": scalar tensor}
model    = {"kind": "speed"|"delta", "Ds","Du", "angle": [..], "not_angle": [..],
            "vel": [..], "pos": [..], "T": float, "use_trig": bool, "norm": [E]}
policy   = {"kind": "plain"|"angles"|"target", "log_ls": [1,Dp], "centers": [nb,Dp], "W": [Du,nb],
            "bias": None|[Du], "u_max": None|float|[Du], "scale": [1,Dp],
            "angle": [..], "non_angle": [..], "target_traj": [H,Ds]}
"""
from __future__ import annotations

import math

import torch

F64 = torch.float64


# --------------------------------------------------------------------------------------------
# kernel functions  (a2 in SURVEY.md §8)
# --------------------------------------------------------------------------------------------
def se_sqdist(se, X1, X2):
    """Lengthscale-weighted squared distance, matmul form.  Stationary_GP.py:65-101."""
    ls = torch.exp(se["log_ls"])
    A = X1[:, se["active"]] / ls
    a2 = (A * A).sum(1, keepdim=True)
    if X2 is None:
        B, b2 = A, a2
    else:
        B = X2[:, se["active"]] / ls
        b2 = (B * B).sum(1, keepdim=True)
    return a2 + b2.t() - 2.0 * (A @ B.t())


def se_cov(se, X1, X2=None):
    """lambda * exp(-dist): no 1/2 in the exponent.  Stationary_GP.py:162-170."""
    return torch.exp(se["log_lambda"]) * torch.exp(-se_sqdist(se, X1, X2))


def _mpk_phi(m, X):
    """Regression matrix [X[:,active], 1?].  Sparse_GP.py:391-399."""
    P = X[:, m["active"]]
    if m["offset"]:
        P = torch.cat([P, torch.ones(X.shape[0], 1, dtype=X.dtype)], 1)
    return P


def _mpk_sigma_diag(m, d):
    """Diagonal of Sigma for factor d of an MPK of degree P: ((P-d)*exp(p_d))**2.

    Sparse_GP.py:613-623 — the loop over ``deg`` ignores its loop variable and adds the *same*
    slice (P-d) times; Utils/Parameters_covariance_functions.py:18-24 squares it on the diagonal.
    """
    n = m["log_par"].numel() // m["deg"]
    acc = torch.zeros(n, dtype=F64)
    for _ in range(d, m["deg"]):
        acc = acc + torch.exp(m["log_par"][d * n:(d + 1) * n])
    return acc ** 2


def mpk_cov(m, X1, X2=None):
    """prod_d phi(X1) diag(Sigma_d) phi(X2)^T.  Sparse_GP.py:625-634, 426-441."""
    P1 = _mpk_phi(m, X1)
    P2 = P1 if X2 is None else _mpk_phi(m, X2)
    K = torch.ones(P1.shape[0], P2.shape[0], dtype=F64)
    for d in range(m["deg"]):
        K = K * (P1 @ (torch.diag(_mpk_sigma_diag(m, d)) @ P2.t()))
    return K


def mpk_diag(m, X):
    """Diagonal of mpk_cov.  Sparse_GP.py:657-668, 443-453."""
    P = _mpk_phi(m, X)
    dg = torch.ones(X.shape[0], dtype=F64)
    for d in range(m["deg"]):
        dg = dg * ((P @ torch.diag(_mpk_sigma_diag(m, d))) * P).sum(1)
    return dg


def sigma_n2(spec):
    """exp(sigma_n_log)**2 + sigma_n_num**2.  GP_prior.py:87-89, 290-296."""
    return torch.exp(spec["sigma_n_log"]) ** 2 + spec["sigma_n_num"] ** 2


def gp_cov(spec, X1, X2=None, noise=False):
    """Sum of the child covariances (+ noise on the diagonal).  GP_prior.py:314-335."""
    parts = []
    if spec["se"] is not None:
        parts.append(se_cov(spec["se"], X1, X2).unsqueeze(0))
    for m in spec["mpk"]:
        parts.append(mpk_cov(m, X1, X2).unsqueeze(0))
    K = torch.cat(parts, 0).sum(0)
    if noise:
        K = K + sigma_n2(spec) * torch.eye(X1.shape[0], dtype=F64)
    return K


def gp_diag(spec, X):
    """k(x,x) without noise.  GP_prior.py:337-347; Stationary_GP.py:172-181."""
    dg = torch.zeros(X.shape[0], dtype=F64)
    if spec["se"] is not None:
        dg = dg + torch.exp(spec["se"]["log_lambda"]) * torch.ones(X.shape[0], dtype=F64)
    for m in spec["mpk"]:
        dg = dg + mpk_diag(m, X)
    return dg


def gp_mean(spec, X):
    """Prior mean: constant of the first child (Sum_Independent_GP.get_mean returns inside the
    loop, GP_prior.py:306-312); MPK children have no mean (Sparse_GP.py:418-424)."""
    if spec["se"] is not None:
        return spec["se"]["mean"].repeat(X.shape[0], 1)
    return torch.zeros(X.shape[0], 1, dtype=F64)


# --------------------------------------------------------------------------------------------
# per-model-update precompute (a1) and posterior (a3)
# --------------------------------------------------------------------------------------------
def gp_fit(spec, X, y):
    """K = k(X,X)+sn2 I; U = chol(K) upper; K^-1 = U^-1 U^-T; alpha = K^-1 (y - m).

    GP_prior.py:91-115 (forward), :130-135 (get_alpha).  Returns (alpha [N,1], m_X [N,1], K_inv [N,N]).
    """
    K = gp_cov(spec, X, None, noise=True)
    U = torch.linalg.cholesky(K, upper=True)
    U_inv = torch.inverse(U)
    K_inv = U_inv @ U_inv.t()
    m_X = gp_mean(spec, X)
    alpha = K_inv @ (y - m_X)
    return alpha, m_X, K_inv


def gp_predict(spec, Xtr, alpha, K_inv, Xs):
    """mean = m(x*) + K* alpha ; var = k** - rowsum((K* K^-1) * K*).  GP_prior.py:137-155."""
    Ks = gp_cov(spec, Xs, Xtr)
    mean = gp_mean(spec, Xs) + Ks @ alpha
    var = gp_diag(spec, Xs) - ((Ks @ K_inv) * Ks).sum(1)
    return mean, var


def sod_select(spec, X, Y, threshold):
    """Greedy subset-of-data selection.  GP_prior.py:232-257 (no permutation)."""
    idx = [0]
    for i in range(1, X.shape[0]):
        alpha, _, K_inv = gp_fit(spec, X[idx, :], Y[idx, :])
        _, var = gp_predict(spec, X[idx, :], alpha, K_inv, X[i:i + 1, :])
        if torch.sqrt(var) > threshold:
            idx.append(i)
    return idx


# --------------------------------------------------------------------------------------------
# one model step (a4)
# --------------------------------------------------------------------------------------------
def gp_features(model, x, u):
    """[x[not_angle], sin x[angle], cos x[angle], u]  (Model_learning.py:670-683, 564-579) or
    [x, u] for the plain model (Model_learning.py:450-456)."""
    if model["use_trig"]:
        ext = torch.cat([x[:, model["not_angle"]], torch.sin(x[:, model["angle"]]),
                         torch.cos(x[:, model["angle"]])], 1)
        return torch.cat([ext, u], 1)
    return torch.cat([x, u], 1)


def next_state(model, gps, x, u, eps, particle_pred=True):
    """GP predict per output, reparameterised sample, integrate.

    Model_learning.py:210-229 (var scaled by norm**2, mean not), :231-242, :315-336;
    speed integration :685-718, delta-state :471-493.  ``gps`` is a list of
    (spec, Xtr, alpha, K_inv); ``eps`` [M,E] is the standard-normal draw of Normal.rsample.
    Returns (x_next, mean [M,E], var [M,E]).
    """
    xt = gp_features(model, x, u)
    means, vars_ = [], []
    for i, (spec, Xtr, alpha, K_inv) in enumerate(gps):
        m, v = gp_predict(spec, Xtr, alpha, K_inv, xt)
        means.append(m)
        vars_.append(v.reshape(-1, 1) * model["norm"][i] ** 2)
    mu = torch.cat(means, 1)
    var = torch.cat(vars_, 1)
    delta = mu + torch.sqrt(var) * eps if particle_pred else mu
    if model["kind"] == "speed":
        nxt = torch.zeros_like(x)
        nxt[:, model["vel"]] = x[:, model["vel"]] + delta
        nxt[:, model["pos"]] = x[:, model["pos"]] + model["T"] * x[:, model["vel"]] + model["T"] / 2 * delta
    else:
        nxt = x + delta
    return nxt, mu, var


# --------------------------------------------------------------------------------------------
# policy (a5)
# --------------------------------------------------------------------------------------------
def policy_features(pol, x, t):
    """Policy.py:323-335 (cos before sin!), :389-403, :242-250."""
    if pol["kind"] == "angles":
        return torch.cat([x[:, pol["non_angle"]], torch.cos(x[:, pol["angle"]]), torch.sin(x[:, pol["angle"]])], 1)
    if pol["kind"] == "target":
        tgt = pol["target_traj"][t, :]
        return torch.cat([x, tgt.repeat(1, x.shape[0]).view(x.shape) - x], 1)
    return x


def policy_apply(pol, x, t, mask, p):
    """Sum of Gaussians + dropout + linear + tanh squashing.  Policy.py:242-265, 52-60.

    ``mask`` [M,nb] in {0,1} stands for the Bernoulli(1-p) draw of F.dropout (training=True), whose
    kept units are scaled by 1/(1-p); ``mask=None`` means no dropout at all (p = 0).
    """
    z = policy_features(pol, x, t).unsqueeze(1) / pol["scale"]
    ls = torch.exp(pol["log_ls"])
    zs = z / ls
    cs = pol["centers"] / ls
    dist = (zs ** 2).sum(2, keepdim=True)
    dist = dist + (cs ** 2).sum(1, keepdim=True).t()
    dist = dist - 2.0 * torch.matmul(zs, cs.t())
    h = torch.exp(-dist)
    if mask is not None and p > 0.0:
        h = h * mask.unsqueeze(1) / (1.0 - p)
    a = torch.matmul(h, pol["W"].t())
    if pol["bias"] is not None:
        a = a + pol["bias"]
    a = a.reshape(-1, pol["W"].shape[0])
    if pol["u_max"] is None:
        return a
    um = torch.as_tensor(pol["u_max"], dtype=F64)
    return um * torch.tanh(a / um)


# --------------------------------------------------------------------------------------------
# costs (a6)
# --------------------------------------------------------------------------------------------
def cost_cart_pole(states, target, ls, angle_index, pos_index):
    """1 - exp(-((|theta|-theta*)/l0)^2 - ((p-p*)/l1)^2).  Cost_function.py:170-182."""
    th = states[:, :, angle_index]
    px = states[:, :, pos_index]
    return 1 - torch.exp(-(((torch.abs(th) - target[0]) / ls[0]) ** 2) - ((px - target[1]) / ls[1]) ** 2)


def _sqdist_target(states, target, ls, active):
    """matmul-form distance to a (set of) target(s).  Cost_function.py:53-63."""
    ns = states[:, :, active] / ls
    nt = target / ls
    d = (ns ** 2).sum(2, keepdim=True)
    d = d + (nt ** 2).sum(1, keepdim=True).t()
    return d - 2.0 * torch.matmul(ns, nt.t())


def cost_distance(states, target, ls, active):
    """Cost_function.py:53-63."""
    return _sqdist_target(states, target, ls, active)


def cost_saturated_distance(states, target, ls, active):
    """1 - exp(-dist).  Cost_function.py:80-101."""
    return 1 - torch.exp(-_sqdist_target(states, target, ls, active))


def cost_saturated_trajectory(states, target_traj, ls, used=None):
    """1 - exp(-sum_j((x_j - target_tj)/l_j)^2).  Cost_function.py:124-147."""
    if used is None:
        used = list(range(states.shape[2]))
    tg = target_traj.repeat(1, states.shape[1]).view(states.shape)
    d = (((states[:, :, used] - tg[:, :, used]) / ls) ** 2).sum(2)
    return 1 - torch.exp(-d)


def expected_cost(costs):
    """sum_t mean_m c ; sum_t std_m(c.detach()) (unbiased).  Cost_function.py:25-36."""
    if costs.dim() == 3:  # [H,M,1] from the matmul-form costs
        costs = costs.squeeze(2) if costs.shape[2] == 1 else costs
    return costs.mean(1).sum(), torch.std(costs.detach(), 1).sum()


# --------------------------------------------------------------------------------------------
# rollouts (a7)
# --------------------------------------------------------------------------------------------
def rollout(model, gps, pol, x0, eps, masks, p_dropout, particle_pred=True):
    """H-step particle rollout.  MC_PILCO.py:659-674.

    x0 [M,Ds] initial particles (the caller draws them: mean + sqrt(var)*eps0, MC_PILCO.py:648-657),
    eps [H-1,M,E], masks [H,M,nb] or None.  Returns states [H,M,Ds], inputs [H,M,Du].
    """
    H = (eps.shape[0] + 1) if eps is not None else masks.shape[0]
    xs = [x0]
    us = [policy_apply(pol, x0, 0, None if masks is None else masks[0], p_dropout)]
    for t in range(1, H):
        nxt, _, _ = next_state(model, gps, xs[t - 1], us[t - 1], eps[t - 1], particle_pred)
        xs.append(nxt)
        us.append(policy_apply(pol, nxt, t, None if masks is None else masks[t], p_dropout))
    return torch.stack(xs), torch.stack(us)


def butter1(fc):
    """First-order Butterworth low-pass (scipy.signal.butter(1, fc)), closed form via the bilinear
    transform: k = tan(pi*fc/2); b = [k, k]/(1+k); a = [1, (k-1)/(k+1)].  MC_PILCO.py:859."""
    k = math.tan(math.pi * fc / 2.0)
    return [k / (1.0 + k), k / (1.0 + k)], [1.0, (k - 1.0) / (k + 1.0)]


def rollout_4pms(model, gps, pol, x0, eps, masks, p_dropout, meas_eps, std_pos, pos_idx, vel_idx, T, fc):
    """Rollout with simulated measurement: position noise, finite-difference velocity and an online
    first-order low-pass; the policy sees the measured state.  MC_PILCO.py:846-906.

    meas_eps [H-1,M,n_pos] stands for torch.randn (MC_PILCO.py:884).
    """
    b, a = butter1(fc)
    H = eps.shape[0] + 1
    xs = [x0]
    noisy = [x0.clone()]
    meas = [noisy[0]]
    us = [policy_apply(pol, meas[0], 0, None if masks is None else masks[0], p_dropout)]
    for t in range(1, H):
        nxt, _, _ = next_state(model, gps, xs[t - 1], us[t - 1], eps[t - 1])
        xs.append(nxt)
        nz = nxt.clone()
        nz[:, pos_idx] = nz[:, pos_idx] + std_pos * meas_eps[t - 1]
        nz[:, vel_idx] = (nz[:, pos_idx] - noisy[t - 1][:, pos_idx]) / T
        noisy.append(nz)
        ms = nz.clone()
        ms[:, vel_idx] = (b[0] * nz[:, vel_idx] + b[1] * noisy[t - 1][:, vel_idx] - a[1] * meas[t - 1][:, vel_idx]) / a[0]
        meas.append(ms)
        us.append(policy_apply(pol, ms, t, None if masks is None else masks[t], p_dropout))
    return torch.stack(xs), torch.stack(us)


def initial_particles(mean, var, eps0):
    """Gaussian initial particles: MultivariateNormal(mean, diag(var)).rsample() = mean + sqrt(var)*eps
    (scale_tril of a diagonal covariance is diag(sqrt(var))).  MC_PILCO.py:648-657."""
    return mean.reshape(1, -1) + torch.sqrt(var).reshape(1, -1) * eps0


# --------------------------------------------------------------------------------------------
# spec helpers shared by tests / bench (synthetic workloads of SURVEY.md §8d)
# --------------------------------------------------------------------------------------------
def make_spec(D, log_ls=None, log_lambda=0.0, mean=0.0, mpk_log_pars=(), sigma_n=0.1, sigma_n_num=0.0):
    """Build a gp spec.  ``mpk_log_pars[k]`` are the log-parameters of the degree-(k+1) MPK term of a
    Volterra series (Sparse_GP.py:671-737): term 0 has an offset column, higher terms do not."""
    act = torch.arange(D)
    se = None
    if log_ls is not None:
        se = {"active": act, "log_ls": torch.as_tensor(log_ls, dtype=F64),
              "log_lambda": torch.tensor([log_lambda], dtype=F64), "mean": torch.tensor([mean], dtype=F64)}
    mpk = []
    for k, lp in enumerate(mpk_log_pars):
        mpk.append({"active": act, "deg": k + 1, "offset": k == 0, "log_par": torch.as_tensor(lp, dtype=F64)})
    return {"D": D, "se": se, "mpk": mpk, "sigma_n_log": torch.tensor(math.log(sigma_n), dtype=F64),
            "sigma_n_num": torch.tensor(sigma_n_num, dtype=F64)}


def cartpole_ode(s, u):
    """Cart-pole dynamics constants of simulation_class/ode_systems.py:43-66 (used only to synthesise
    smooth training targets)."""
    mc, mp, ll, g, bb = 0.5, 0.5, 0.5, 9.81, 0.1
    p, dp, th, dth = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
    st, ct = torch.sin(th), torch.cos(th)
    den = 4 * (mc + mp) - 3 * mp * ct ** 2
    ddp = (2 * mp * ll * dth ** 2 * st + 3 * mp * g * st * ct + 4 * u - 4 * bb * dp) / den
    ddth = (-3 * mp * ll * dth ** 2 * st * ct - 6 * (mc + mp) * g * st - 6 * (u - bb * dp) * ct) / (ll * den)
    return torch.stack([dp, ddp, dth, ddth], 1)


def cartpole_dataset(N, sigma_n, gen):
    """Synthetic cart-pole transitions: gp inputs [p, dp, dtheta, sin, cos, u] and the two velocity
    increments after one RK4 step of 0.05 s."""
    r = torch.rand(N, 5, dtype=F64, generator=gen)
    s = torch.stack([4 * r[:, 0] - 2, 10 * r[:, 1] - 5, 2 * math.pi * r[:, 2] - math.pi, 20 * r[:, 3] - 10], 1)
    u = 20 * r[:, 4] - 10
    dt = 0.05
    k1 = cartpole_ode(s, u); k2 = cartpole_ode(s + dt / 2 * k1, u)
    k3 = cartpole_ode(s + dt / 2 * k2, u); k4 = cartpole_ode(s + dt * k3, u)
    s1 = s + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    X = torch.stack([s[:, 0], s[:, 1], s[:, 3], torch.sin(s[:, 2]), torch.cos(s[:, 2]), u], 1)
    nz = sigma_n * torch.randn(N, 2, dtype=F64, generator=gen)
    Y = torch.stack([s1[:, 1] - s[:, 1], s1[:, 3] - s[:, 3]], 1) + nz
    return X, Y


# --------------------------------------------------------------------------------------------
# hyper-parameter training objective
# --------------------------------------------------------------------------------------------
def nlml(spec, X, y):
    """0.5 ((y - m)^T K^-1 (y - m) + log det K), no N log 2 pi term.  Likelihood/Gaussian_likelihood.py:12-24 on the outputs of
    GP_prior.forward (GP_prior.py:91-115: upper Cholesky, explicit inverse, log det from the factor's diagonal).
    Differentiable w.r.t. any spec tensor that requires grad (the reference trains through this very graph)."""
    K = gp_cov(spec, X, None, noise=True)
    U = torch.linalg.cholesky(K, upper=True)
    log_det = 2 * torch.sum(torch.log(torch.diag(U)))
    U_inv = torch.inverse(U)
    K_inv = U_inv @ U_inv.t()
    r = y - gp_mean(spec, X)
    return 0.5 * (r.t() @ (K_inv @ r) + log_det)

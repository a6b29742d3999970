"""Host-side flattening of the reference's object trees into the POD structs of include/mcpilco_b200.h.

Pure host logic (numpy + ctypes): nothing here touches CUDA, so it is unit-tested on CPU.  The
parameterisation follows the reference exactly (file:line relative to the reference root):

  * SE:   lambda = exp(log_lambda_par), 1/l_j = exp(-log_lengthscales_par[j])   Stationary_GP.py:86-101,162-170
  * MPK:  factor d of a degree-P term has diagonal Sigma ((P-d) * exp(p_d))**2   Sparse_GP.py:613-623
          (the loop adds the SAME slice P-d times), with an extra offset column when flg_offset
          (Sparse_GP.py:391-399); Utils/Parameters_covariance_functions.py:18-24 squares the diagonal.
  * noise sigma_n2 = exp(sigma_n_log)**2 + sigma_n_num**2                        GP_prior.py:87-89
"""
import ctypes as C

import numpy as np

from . import _native as N


def _np(x):
    """numpy float64 view of a tensor / array / scalar (host copy if it lives on a device)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def new_gp_spec(D):
    if not 1 <= D <= N.MAX_D:
        raise ValueError("gp input dimension %d outside [1, %d]" % (D, N.MAX_D))
    s = N.GpSpec()
    s.D = D
    s.lambda_ = 1.0
    for p in range(N.MAX_POLY):
        for f in range(N.MAX_DEG):
            s.poly_w2[p][f][N.MAX_D] = 1.0  # neutral factor: 0 * <x,y> + 1
    return s


def add_se(spec, active, log_ls, log_lambda, mean):
    """RBF child (Stationary_GP.py:112-181).  A scalar lengthscale (flg_ARD False) broadcasts."""
    if spec.has_se:
        raise ValueError("only one squared-exponential term per GP is supported")
    active = np.asarray(active, dtype=np.int64).reshape(-1)
    inv = np.exp(-_np(log_ls)).reshape(-1)
    if inv.size == 1:
        inv = np.repeat(inv, active.size)
    if inv.size != active.size or (active.size and (active.min() < 0 or active.max() >= spec.D)):
        raise ValueError("SE lengthscales / active_dims do not match the gp input dimension")
    spec.has_se = 1
    spec.lambda_ = float(np.exp(_np(log_lambda)).reshape(-1)[0])
    for j in range(N.MAX_D):
        spec.inv_ls[j] = 0.0
    for a, v in zip(active, inv):
        spec.inv_ls[int(a)] = float(v)
    return spec


def add_mpk(spec, active, deg, offset, log_par):
    """One multiplicative-polynomial term of degree `deg` (Sparse_GP.py:559-668)."""
    active = np.asarray(active, dtype=np.int64).reshape(-1)
    p = spec.n_poly
    if p >= N.MAX_POLY:
        raise ValueError("more than %d polynomial terms" % N.MAX_POLY)
    if not 1 <= deg <= N.MAX_DEG:
        raise ValueError("polynomial degree %d outside [1, %d]" % (deg, N.MAX_DEG))
    lp = _np(log_par).reshape(-1)
    n = active.size + (1 if offset else 0)
    if lp.size != deg * n or (active.size and (active.min() < 0 or active.max() >= spec.D)):
        raise ValueError("MPK parameters: expected %d values, got %d" % (deg * n, lp.size))
    for f in range(N.MAX_DEG):
        for j in range(N.MAX_D):
            spec.poly_w2[p][f][j] = 0.0
        spec.poly_w2[p][f][N.MAX_D] = 1.0
    for d in range(deg):
        w2 = ((deg - d) * np.exp(lp[d * n:(d + 1) * n])) ** 2
        for a, v in zip(active, w2[:active.size]):
            spec.poly_w2[p][d][int(a)] = float(v)
        spec.poly_w2[p][d][N.MAX_D] = float(w2[active.size]) if offset else 0.0
    spec.poly_deg[p] = deg
    spec.n_poly = p + 1
    return spec


def spec_from_dict(d):
    """Flat description used by tests/bench:  {"D", "log_ls"|None, "lambda", "mean", "mpk": [w...], "sigma_n"}
    where mpk[k] holds the (positive) Sigma_pos_par_init of the degree-(k+1) Volterra term
    (first term with offset, Sparse_GP.py:671-737)."""
    D = int(d["D"])
    s = new_gp_spec(D)
    act = np.arange(D)
    if d.get("log_ls") is not None:
        add_se(s, act, d["log_ls"], np.log(d.get("lambda", 1.0)), d.get("mean", 0.0))
        s.mean0 = float(d.get("mean", 0.0))
    for k, w in enumerate(d.get("mpk", [])):
        add_mpk(s, act, k + 1, k == 0, np.log(np.asarray(w, dtype=np.float64)))
    s.sigma_n2 = float(d.get("sigma_n", 0.0)) ** 2 + float(d.get("sigma_n_num", 0.0)) ** 2
    return s


def _fill_idx(dst, values, cap, what):
    values = [int(v) for v in np.asarray(values).reshape(-1)]
    if len(values) > cap:
        raise ValueError("%s: %d entries exceed the ABI limit %d" % (what, len(values), cap))
    for i, v in enumerate(values):
        dst[i] = v
    return len(values)


def model_struct(kind, Ds, Du, E, angle=(), not_angle=(), vel=(), pos=(), T=0.0, use_trig=None, particle_pred=True):
    """kind: "speed" (Model_learning.py:685-718) or "delta" (:471-493)."""
    m = N.Model()
    if use_trig is None:
        use_trig = len(angle) > 0 or kind == "speed"
    m.Ds, m.Du, m.E = int(Ds), int(Du), int(E)
    m.kind = 1 if kind == "speed" else 0
    m.use_trig = 1 if use_trig else 0
    m.n_na = _fill_idx(m.na_idx, not_angle, N.MAX_DS, "not_angle_indeces") if use_trig else 0
    m.n_a = _fill_idx(m.a_idx, angle, N.MAX_DS, "angle_indeces") if use_trig else 0
    m.D = (m.n_na + 2 * m.n_a + m.Du) if use_trig else (m.Ds + m.Du)
    if m.kind == 1:
        nv = _fill_idx(m.vel_idx, vel, N.MAX_E, "vel_indeces")
        npos = _fill_idx(m.pos_idx, pos, N.MAX_E, "not_vel_indeces")
        if nv != E or npos != E:
            raise ValueError("speed model: need one velocity and one position index per GP")
    m.particle_pred = 1 if particle_pred else 0
    m.T = float(T)
    return m


def policy_struct(kind, nb, Dp, Du, Ds, u_max=None, scale=None, angle=(), non_angle=(), has_bias=False, use_drop=True):
    """kind: "plain" (Policy.py:242-265), "angles" (:323-335), "target" (:389-403)."""
    p = N.Policy()
    p.kind = {"plain": 0, "angles": 1, "target": 2}[kind]
    p.nb, p.Dp, p.Du, p.Ds = int(nb), int(Dp), int(Du), int(Ds)
    if Dp > N.MAX_DP or Du > N.MAX_DU or Ds > N.MAX_DS:
        raise ValueError("policy dimensions exceed the ABI limits")
    if kind == "angles":
        p.n_na = _fill_idx(p.na_idx, non_angle, N.MAX_DS, "non_angle_indices")
        p.n_a = _fill_idx(p.a_idx, angle, N.MAX_DS, "angle_indices")
    p.squash = 0 if u_max is None else 1
    um = np.ones(Du) if u_max is None else np.broadcast_to(np.asarray(u_max, dtype=np.float64).reshape(-1), (Du,))
    for k in range(Du):
        p.u_max[k] = float(um[k])
    sc = np.ones(Dp) if scale is None else _np(scale).reshape(-1)
    for j in range(Dp):
        p.inv_scale[j] = 1.0 / float(sc[j])
    p.has_bias = 1 if has_bias else 0
    p.use_drop = 1 if use_drop else 0
    return p


def cost_struct(kind, Ds, **kw):
    """kind: None | "cart_pole" | "sat_traj" | "sat_target" | "distance" (Cost_function.py:53-182)."""
    c = N.Cost()
    if kind is None:
        c.kind = 0
        return c
    if kind == "cart_pole":
        c.kind = 1
        c.n_idx = 2
        c.idx[0], c.idx[1] = int(kw["angle_index"]), int(kw["pos_index"])
        tg, ls = _np(kw["target"]).reshape(-1), _np(kw["ls"]).reshape(-1)
        for i in range(2):
            c.target[i] = float(tg[i])
            c.inv_ls[i] = 1.0 / float(ls[i])
        return c
    if kind == "sat_traj":
        c.kind = 2
        used = kw.get("used")
        used = list(range(Ds)) if used is None else list(used)
        ls = np.broadcast_to(_np(kw["ls"]).reshape(-1), (len(used),))
    else:
        c.kind = 3 if kind == "sat_target" else 4
        used = list(kw["active"])
        tg = _np(kw["target"])
        if tg.ndim == 2 and tg.shape[0] != 1:
            raise ValueError("fused cost supports a single target state")
        tg = tg.reshape(-1)
        ls = np.broadcast_to(_np(kw["ls"]).reshape(-1), (len(used),))
        for i in range(len(used)):
            c.target[i] = float(tg[i])
    c.n_idx = _fill_idx(c.idx, used, N.MAX_DS, "cost dims")
    for i in range(len(used)):
        c.inv_ls[i] = 1.0 / float(ls[i])
    return c


def meas_struct(pos_idx=None, vel_idx=None, std_pos=None, fc=None, T=0.0):
    """MC_PILCO4PMS measurement model; (b, a) = scipy.signal.butter(1, fc) in closed form
    (bilinear transform of a first-order low-pass: k = tan(pi fc / 2))."""
    m = N.Meas()
    if pos_idx is None:
        return m
    m.enabled = 1
    m.n_pos = _fill_idx(m.pos_idx, pos_idx, N.MAX_E, "pos_indeces")
    if _fill_idx(m.vel_idx, vel_idx, N.MAX_E, "vel_indeces") != m.n_pos:
        raise ValueError("pos_indeces and vel_indeces must have the same length")
    sp = np.broadcast_to(_np(std_pos).reshape(-1), (m.n_pos,))
    for i in range(m.n_pos):
        m.std_pos[i] = float(sp[i])
    k = np.tan(np.pi * float(fc) / 2.0)
    m.b0 = m.b1 = k / (1.0 + k)
    m.a0, m.a1 = 1.0, (k - 1.0) / (k + 1.0)
    m.T = float(T)
    return m


def struct_bytes(s):
    return bytes(memoryview(s).cast("B")) if not isinstance(s, C.Structure) else C.string_at(C.addressof(s), C.sizeof(s))

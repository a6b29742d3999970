"""Expected-cost objects — the reference's `policy_learning/Cost_function.py` surface (Expected_cost :15-36 and the concrete
costs :39-182).  cost = sum_t mean_m c(x_t^m) ; std_cost = sum_t std_m c (unbiased, on detached costs).

Two paths, chosen per call:
  * fused: when `states_sequence` is the tensor `MC_PILCO.apply_policy` just returned and this object is the rollout's cost,
    the per-particle costs, their mean/std and — in backward — their gradient were/are computed inside the CUDA rollout
    (no [H, M] temporaries, no autograd graph); the two scalars returned are outputs of the rollout's autograd node.
  * generic: any other CUDA `states_sequence` (or a user-supplied `Expected_cost(cost_function=...)` lambda) is evaluated with
    torch ops on the device; autograd then hands d cost / d states to the hand-written backward kernel.
"""
import torch

from .. import _pack as P
from .. import distributed as D


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("mcpilco_b200 cost functions take CUDA tensors (this path has no CPU fallback)")


class Expected_cost(torch.nn.modules.loss._Loss):
    def __init__(self, cost_function):
        super().__init__()
        self.cost_function = cost_function

    def fused_spec(self, Ds, H, trial_index=None):
        """(McpCost, target trajectory tensor or None) when the CUDA rollout can evaluate this cost itself, else None."""
        return None

    def fused_spec_cached(self, Ds, H, trial_index=None):
        """fused_spec memoised on the object's tensor attributes (flattening reads them back to the host: not once per rollout)."""
        key = (Ds, H, trial_index) + tuple((k, id(v), v._version, v.data_ptr()) for k, v in sorted(vars(self).items()) if isinstance(v, torch.Tensor))
        hit = getattr(self, "_fused_memo", None)
        if hit is None or hit[0] != key:
            hit = (key, self.fused_spec(Ds, H, trial_index))
            self._fused_memo = hit
        return hit[1]

    def forward(self, states_sequence, inputs_sequence, trial_index=None):
        fused = getattr(states_sequence, "_mcp_fused_cost", None)
        if fused is not None and fused[0] is self and (fused[1] is None or fused[1] == trial_index):
            return fused[2], fused[3]
        _need_cuda(states_sequence)
        costs = self.cost_function(states_sequence, inputs_sequence, trial_index)
        shard = getattr(states_sequence, "_mcp_shard", None)
        if shard is None:
            return torch.sum(torch.mean(costs, 1)), torch.sum(torch.std(costs.detach(), 1))
        # particles sharded over ranks: merge the per-step moments so that every rank sees the GLOBAL cost / std (and takes the same
        # control-flow decisions in reinforce_policy); the gradient of the global particle mean w.r.t. this shard's costs is 1 / M_global
        rank, world, group, m_global = shard
        H, n_local = costs.shape[0], costs.shape[1]
        rows = costs.reshape(H, n_local, -1).permute(0, 2, 1).reshape(-1, n_local)  # [H * k, M_local]: one row per (step, cost column)
        local_mean = rows.mean(1)
        counts = [D.shard(m_global, r, world)[1] for r in range(world)]
        if counts[rank] != n_local:
            raise RuntimeError("Expected_cost: states hold %d particles but this rank's shard has %d" % (n_local, counts[rank]))
        with torch.no_grad():
            stats = torch.stack([local_mean, ((rows - local_mean.unsqueeze(1)) ** 2).sum(1)], 1)
            mean, m2 = D.merge_cost_stats(D.gather_cost_stats(stats, group, world), counts)
            cost_g, std_g = D.expected_cost_from_stats(mean, m2, m_global)
        local = local_mean.sum() * (n_local / float(m_global))
        return local + (cost_g - local.detach()), std_g


def _on(v, like):
    return torch.as_tensor(v, dtype=like.dtype, device=like.device)


def _sqdist(states_sequence, target_state, lengthscales, active_dims):
    lengthscales = _on(lengthscales, states_sequence)
    ns = states_sequence[:, :, active_dims] / lengthscales
    nt = _on(target_state, states_sequence) / lengthscales
    return ((ns.unsqueeze(2) - nt.reshape(1, 1, -1, ns.shape[2])) ** 2).sum(3)


def distance_from_target(states_sequence, inputs_sequence, trial_index, target_state, lengthscales, active_dims):
    """Lengthscale-weighted squared distance to each target row: [H, M, n_targets] (reference :53-63)."""
    return _sqdist(states_sequence, target_state, lengthscales, active_dims)


def saturated_distance_from_target(states_sequence, inputs_sequence, trial_index, target_state, lengthscales, active_dims):
    """1 - exp(-distance) (reference :80-101)."""
    return 1 - torch.exp(-_sqdist(states_sequence, target_state, lengthscales, active_dims))


def saturated_distance_from_trajectory(states_sequence, inputs_sequence, trial_index, target_traj, lengthscales, flg_var_lengthscales,
                                       used_indeces):
    """1 - exp(-sum_j ((x_tj - target_tj) / l_j)^2) (reference :124-147)."""
    if used_indeces is None:
        used_indeces = list(range(states_sequence.shape[2]))
    ls = _on(lengthscales[trial_index] if flg_var_lengthscales else lengthscales, states_sequence)
    tg = _on(target_traj, states_sequence)[:states_sequence.shape[0], :].unsqueeze(1)
    d = (((states_sequence[:, :, used_indeces] - tg[:, :, used_indeces]) / ls) ** 2).sum(2)
    return 1 - torch.exp(-d)


def cart_pole_cost(states_sequence, inputs_sequence, trial_index, target_state, lengthscales, angle_index, pos_index):
    """1 - exp(-((|theta| - theta*) / l0)^2 - ((p - p*) / l1)^2) (reference :170-182)."""
    x, theta = states_sequence[:, :, pos_index], states_sequence[:, :, angle_index]
    target_state, lengthscales = _on(target_state, states_sequence), _on(lengthscales, states_sequence)
    return 1 - torch.exp(-(((torch.abs(theta) - target_state[0]) / lengthscales[0]) ** 2) - ((x - target_state[1]) / lengthscales[1]) ** 2)


class _Target_cost(Expected_cost):
    _kind = None

    def __init__(self, target_state, lengthscales, active_dims):
        self.target_state, self.lengthscales, self.active_dims = target_state, lengthscales, list(active_dims)
        fn = distance_from_target if self._kind == "distance" else saturated_distance_from_target
        super().__init__(lambda x, u, trial_index: fn(x, u, trial_index, target_state=target_state, lengthscales=lengthscales,
                                                      active_dims=self.active_dims))

    def fused_spec(self, Ds, H, trial_index=None):
        if torch.as_tensor(self.target_state).numel() != len(self.active_dims):
            return None  # several target rows: generic path
        return P.cost_struct(self._kind, Ds, target=self.target_state, ls=self.lengthscales, active=self.active_dims), None


class Expected_distance(_Target_cost):
    """Sum of expected squared distances from a target state (reference :39-50)."""
    _kind = "distance"


class Expected_saturated_distance(_Target_cost):
    """Sum of expected saturated distances from a target state (reference :66-77)."""
    _kind = "sat_target"


class Expected_saturated_distance_from_trajectory(Expected_cost):
    """Sum of expected saturated distances from a target trajectory (reference :104-121)."""

    def __init__(self, target_traj, lengthscales, flg_var_lengthscales=False, used_indeces=None):
        self.target_traj, self.lengthscales = target_traj, lengthscales
        self.flg_var_lengthscales, self.used_indeces = flg_var_lengthscales, used_indeces
        super().__init__(lambda x, u, trial_index: saturated_distance_from_trajectory(
            x, u, trial_index, target_traj=target_traj, lengthscales=lengthscales, flg_var_lengthscales=flg_var_lengthscales,
            used_indeces=used_indeces))

    def fused_spec(self, Ds, H, trial_index=None):
        if self.flg_var_lengthscales and trial_index is None:
            return None
        ls = self.lengthscales[trial_index] if self.flg_var_lengthscales else self.lengthscales
        return P.cost_struct("sat_traj", Ds, ls=ls, used=self.used_indeces), self.target_traj


class Cart_pole_cost(Expected_cost):
    """Cart-pole swing-up cost on |theta| and the cart position (reference :150-182)."""

    def __init__(self, target_state, lengthscales, angle_index, pos_index):
        self.target_state, self.lengthscales = target_state, lengthscales
        self.angle_index, self.pos_index = angle_index, pos_index
        super().__init__(lambda x, u, trial_index: cart_pole_cost(x, u, trial_index, target_state=target_state, lengthscales=lengthscales,
                                                                  angle_index=angle_index, pos_index=pos_index))

    def fused_spec(self, Ds, H, trial_index=None):
        return P.cost_struct("cart_pole", Ds, target=self.target_state, ls=self.lengthscales, angle_index=self.angle_index,
                             pos_index=self.pos_index), None

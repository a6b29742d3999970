"""A scripted stand-in for the rollout so that the CONTROL LOGIC of MC_PILCO.reinforce_policy (reference MC_PILCO.py:375-613: NaN
re-sampling, policy re-initialisation and restart, exponential cost-difference monitors, learning-rate halving, dropout reduction,
early exit) can be compared between the reference and mcpilco_b200 on the CPU, without any GP or CUDA.

`install(obj, costs)` replaces obj.apply_policy / obj.cost_function: the k-th rollout's cost is costs[k] * (1 + sum of the policy
parameters * 1e-3) so that backward() produces a gradient and the optimiser really moves the parameters; NaN entries in `costs`
trigger the NaN handling.  Everything is deterministic."""
import numpy as np
import torch


class TinyPolicy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.5, -0.25], dtype=torch.float64))
        self.reinits = 0

    def reinit(self, **kw):
        self.reinits += 1
        self.w.data = torch.tensor([0.1 * self.reinits, -0.1], dtype=torch.float64)


def install(obj, costs, log):
    state = {"k": 0}

    def apply_policy(**kw):
        log.append(("rollout", state["k"], float(kw["p_dropout"])))
        k = state["k"]
        state["k"] += 1
        c = costs[min(k, len(costs) - 1)]
        val = torch.tensor(c, dtype=torch.float64) * (1.0 + 1e-3 * obj.control_policy.w.sum())
        st = val.reshape(1, 1, 1).expand(2, 3, 1)  # [H, M, Ds] carrier of the scripted cost
        return st, torch.zeros(2, 3, 1, dtype=torch.float64)

    def cost_function(states, inputs, trial_index=None):
        return states[0, 0, 0], torch.tensor(0.1, dtype=torch.float64)

    obj.apply_policy = apply_policy
    obj.cost_function = cost_function
    obj.control_policy = TinyPolicy()
    return obj


def script(name):
    rs = np.random.RandomState(5)
    if name == "plateau":  # fast decrease, then a plateau: lr halvings, dropout reduction, early exit
        base = np.concatenate([50.0 * np.exp(-np.arange(60) / 10.0) + 5.0, 5.0 + 0.01 * rs.randn(400)])
        kw = dict(opt_steps_list=[400], lr_list=[0.1], p_dropout_list=[0.25], p_drop_reduction=0.125, min_diff_cost=0.2, num_min_diff_cost=20,
                  min_step=30, lr_min=0.025, lr_reduction_ratio=0.5, alpha_diff_cost=0.9)
    elif name == "nan_retry":  # isolated NaNs are re-sampled, no re-initialisation
        base = 20.0 - 0.05 * np.arange(80)
        base[[7, 8, 30]] = np.nan
        kw = dict(opt_steps_list=[40], lr_list=[0.01], p_dropout_list=[0.1], min_step=np.inf)
    elif name == "nan_reinit":  # ten NaNs in a row: re-initialise the policy and restart the optimisation
        base = np.concatenate([20.0 - 0.05 * np.arange(6), np.full(10, np.nan), 15.0 - 0.05 * np.arange(80)])
        kw = dict(opt_steps_list=[25], lr_list=[0.01], p_dropout_list=None, min_step=np.inf)
    elif name == "nan_at_init":  # NaN in the very first (filter-initialisation) rollouts
        base = np.concatenate([np.full(3, np.nan), 10.0 - 0.1 * np.arange(40)])
        kw = dict(opt_steps_list=[15], lr_list=[0.02], p_dropout_list=[0.0], min_step=np.inf)
    else:
        raise KeyError(name)
    return base, kw


SCRIPTS = ("plateau", "nan_retry", "nan_reinit", "nan_at_init")


def run(obj, name):
    """Run reinforce_policy on the scripted costs; return what the caller and the logs can observe."""
    import contextlib, io
    base, kw = script(name)
    log = []
    install(obj, list(base), log)
    T = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    with contextlib.redirect_stdout(io.StringIO()):
        cost_list, std_list, states, inputs = obj.reinforce_policy(
            T_control=2 * obj.T_sampling, num_particles=3, trial_index=0, particles_initial_state_mean=T([0.0]), particles_initial_state_var=T([1.0]),
            flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None, flg_particles_init_multi_gauss=False,
            f_optimizer="lambda p, lr : torch.optim.Adam(p, lr)", num_step_print=10 ** 9,
            policy_reinit_dict={}, **kw)
    return {"cost_list": np.asarray(cost_list), "std_list": np.asarray(std_list), "n_rollouts": np.array(len(log)),
            "dropouts": np.array([d for _, _, d in log]), "w_final": obj.control_policy.w.detach().numpy().copy(),
            "reinits": np.array(obj.control_policy.reinits), "states": np.asarray(states)}

"""MC-PILCO policy-learning objects — the hot-path part of the reference's `policy_learning/MC_PILCO.py`:
`MC_PILCO.apply_policy` (:615-674), `MC_PILCO.reinforce_policy` (:375-613), `MC_PILCO.rollout` (:347-373) and
`MC_PILCO4PMS.apply_policy` (:808-906), with the reference's constructor and method signatures.

`apply_policy` is ONE fused CUDA rollout (libmcpilco_b200.so) instead of a Python loop of torch ops, and it is ONE autograd
node: `cost.backward()` runs the hand-written backprop-through-time kernel and deposits `.grad` on
`control_policy.log_lengthscales / centers / f_linear.weight (/ bias)` exactly where the reference's autograd graph would.
Under `torch.distributed` (one process per GPU) particles are sharded; cost statistics are all-gathered and the policy
gradient all-reduced inside the same node (mcpilco_b200.distributed).

Out of scope: the trial loop `reinforce`, data collection from the (simulated / real) system, logging
and log re-loading, `MC_PILCO_Experiment`.  Those are host orchestration the reference keeps doing; see INTEGRATION.md for
how its classes bind to these methods.
"""
import time

import numpy as np
import torch

from .. import _ops as ops
from .. import _pack as P
from .. import distributed as D


class _ParticleRollout(torch.autograd.Function):
    """states, inputs, cost, std_cost = rollout(policy parameters).  The graph of the reference (~H*E*40 nodes retaining
    [M, N] tensors, MC_PILCO.py:522) collapses into this single node with O(M*H*E*D) checkpoints.  It sits on the two torch.library
    operators `torch.ops.mcpilco.rollout_fwd` / `rollout_bwd` (torch_ops.RolloutCall)."""

    @staticmethod
    def forward(ctx, call, need_grad, x0, shard, *params):
        states, inputs, cost_out, cost_stats, jac, pol_in = call.forward(x0, need_grad)
        # the trajectories and checkpoints the backward kernel needs travel through autograd's saved-tensor mechanism
        ctx.save_for_backward(states, inputs, jac, pol_in, x0.detach())
        ctx.call, ctx.shard, ctx.need_grad = call, shard, bool(need_grad)
        ctx.want_gx0 = bool(x0.requires_grad)
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        rank, world, group, m_global = shard
        if cost_out is None:
            cost = std = torch.zeros((), dtype=states.dtype, device=states.device)
        elif world == 1:
            cost, std = cost_out[0].clone(), cost_out[1].clone()
        else:
            counts = [D.shard(m_global, r, world)[1] for r in range(world)]
            mean, m2 = D.merge_cost_stats(D.gather_cost_stats(cost_stats, group, world), counts)
            cost, std = D.expected_cost_from_stats(mean, m2, m_global)
        ctx.mark_non_differentiable(std)
        return states, inputs, cost, std

    @staticmethod
    def backward(ctx, g_states, g_inputs, g_cost, g_std):
        call = ctx.call
        if not ctx.need_grad:
            raise RuntimeError("rollout backward: the forward pass was run without need_grad")
        states, inputs, jac, pol_in, x0 = ctx.saved_tensors
        rank, world, group, m_global = ctx.shard
        w_local = call.M / float(m_global)  # this shard's weight in the global particle mean
        generic = g_states is not None or g_inputs is not None
        if generic:
            # g_states / g_inputs are d loss / d (this shard's trajectories) of the GLOBAL loss: Expected_cost weights its local
            # particle mean by w_local itself when sharded; the fused cost's share (if the loss also uses it) is added in the kernel
            gc = 0.0 if g_cost is None or not call.has_cost else float(g_cost) * w_local
            gr = call.backward(x0, states, inputs, jac, pol_in, grad_cost=gc, grad_states=g_states, grad_inputs=g_inputs, want_gx0=ctx.want_gx0)
            scale = None
        else:
            if g_cost is None:
                return (None,) * (4 + ctx.n_params)
            gr = call.backward(x0, states, inputs, jac, pol_in, grad_cost=w_local, want_gx0=ctx.want_gx0)
            scale = g_cost  # stays on the device: no host sync
        keys = ["log_ls", "centers", "W"] + (["bias"] if ctx.n_params == 4 else [])
        if world > 1:
            flat = torch.cat([gr[k].reshape(-1) for k in keys])
            D.allreduce_sum_(flat, group)
            off = 0
            for k in keys:
                n = gr[k].numel()
                gr[k] = flat[off:off + n].view_as(gr[k])
                off += n
        out = [gr[k] if scale is None else gr[k] * scale for k in keys]
        gx0 = gr["x0"] if ctx.want_gx0 else None
        if gx0 is not None and scale is not None:
            gx0 = gx0 * scale
        return (None, None, gx0, None) + tuple(out)


GOLDEN = 0x9E3779B97F4A7C15  # rollout counter -> Philox key stride
MASK64 = (1 << 64) - 1


class _GraphedRollout:
    """One particle rollout forward + hand-written backward, captured ONCE in a CUDA graph and replayed per optimisation step
    (the reference's step is MC_PILCO.py:484-522).  At the real configuration sizes a rollout is ~130-600 dependent
    kernel launches of a few microseconds each: enqueueing them from Python costs more than running them, so the host — not the GPU
    — bounds the optimisation step.  A replay is one cudaGraphLaunch.  What stays on the host is exactly what the reference's loop
    does per step: the NaN test of the cost (its one synchronisation), the monitors and the torch optimiser step.

    Fresh noise per replay without re-capturing: the kernels add a device word to the baked Philox seed (McpNoise.seed_dev); the host
    sets it to GOLDEN * rollout_counter before each replay, which reproduces the key sequence of the un-graphed path bit for bit."""

    def __init__(self, pilco, init, p_dropout):
        self.pilco = pilco
        dev = pilco.device
        pol = pilco.control_policy
        self.params = [pol.log_lengthscales, pol.centers, pol.f_linear.weight] + ([pol.f_linear.bias] if pol.flg_bias else [])
        self.keys = ["log_ls", "centers", "W"] + (["bias"] if pol.flg_bias else [])
        self.signature = self.signature_of(pilco, init, p_dropout)
        as_dev = lambda v: None if v is None else torch.as_tensor(v, dtype=pilco.dtype, device=dev).clone()  # noqa: E731
        # everything that would be a host -> device copy inside the captured region is staged here
        self.init = dict(init)
        for k in ("particles_initial_state_mean", "particles_initial_state_var", "particles_init_up_bound", "particles_init_low_bound"):
            self.init[k] = as_dev(init[k])
        self.p_dropout = float(p_dropout)
        self.ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        if pilco._seed_base is None:
            pilco._next_seed()
            pilco._rollouts -= 1
        self.base = pilco._seed_base & MASK64
        self.call = None
        # eager pass on the capture stream first: first-use initialisation of the native side (kernel attributes, the side streams and
        # scratch that are keyed by the launching stream) must not happen inside the captured region
        self.stream = torch.cuda.Stream(dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            self._run()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.states, self.inputs, self.cost_out, self.grads = self._run()

    @staticmethod
    def signature_of(pilco, init, p_dropout):
        pol, ml = pilco.control_policy, pilco.model_learning
        params = [pol.log_lengthscales, pol.centers, pol.f_linear.weight] + ([pol.f_linear.bias] if pol.flg_bias else [])
        return (float(p_dropout), int(init["num_particles"]), int(init["T_control"]), bool(init["flg_particles_init_uniform"]),
                bool(init["flg_particles_init_multi_gauss"]), tuple(p.data_ptr() for p in params), id(ml.fitted_gps()), pilco._trial_index,
                id(pilco.cost_function))

    @staticmethod
    def eligible(pilco):
        import os
        if os.environ.get("MCPILCO_NO_GRAPH", "0") == "1" or D.world()[1] > 1 or not torch.is_grad_enabled():
            return False
        cf = pilco.cost_function
        return hasattr(cf, "fused_spec") and cf.fused_spec(pilco.state_dim, 2, pilco._trial_index) is not None

    def _run(self):
        p, i = self.pilco, self.init
        H, M = int(i["T_control"]), int(i["num_particles"])
        x0 = p._initial_particles(i["particles_initial_state_mean"], i["particles_initial_state_var"], i["flg_particles_init_uniform"],
                                  i["particles_init_up_bound"], i["particles_init_low_bound"], i["flg_particles_init_multi_gauss"], M, 0,
                                  self.base, None, seed_dev=self.ctr)
        if self.call is None:
            # host-side flattening of model / policy / cost (reads a few device scalars): once, in the eager pass, never while capturing
            self.call = p._rollout_call(x0, M, 0, H, self.p_dropout, self.base, None, M, seed_dev=self.ctr)
        states, inputs, cost_out, _, jac, pol_in = self.call.forward(x0, True)
        grads = self.call.backward(x0, states, inputs, jac, pol_in, grad_cost=1.0)
        return states, inputs, cost_out, grads

    def replay(self):
        """New rollout (next Philox key) + backward; afterwards cost_out / states / inputs / grads hold its results."""
        p = self.pilco
        p._rollouts += 1
        word = (GOLDEN * p._rollouts) & MASK64
        self.ctr.fill_(word - (1 << 64) if word >= (1 << 63) else word)  # the same 64 bits, as the int64 torch stores
        self.graph.replay()
        return self.cost_out[0], self.cost_out[1]

    def deposit_grads(self):
        for prm, k in zip(self.params, self.keys):
            prm.grad = self.grads[k].view_as(prm)


class MC_PILCO(torch.nn.Module):
    """Monte-Carlo Probabilistic Inference for Learning COntrol — hot-path methods (reference :29-751)."""

    def __init__(self, T_sampling, state_dim, input_dim, f_sim, f_model_learning, model_learning_par, f_rand_exploration_policy,
                 rand_exploration_policy_par, f_control_policy, control_policy_par, f_cost_function, cost_function_par,
                 std_meas_noise=None, log_path=None, dtype=torch.float64, device=torch.device("cuda")):
        super().__init__()
        self.T_sampling, self.dtype, self.device = T_sampling, dtype, device
        self.state_dim, self.input_dim = state_dim, input_dim
        self.f_sim = f_sim  # the simulated / real system lives outside this path
        self.std_meas_noise = np.zeros(state_dim) if std_meas_noise is None else std_meas_noise
        self.model_learning = f_model_learning(**model_learning_par)
        self.rand_exploration_policy = None if f_rand_exploration_policy is None else f_rand_exploration_policy(**rand_exploration_policy_par)
        self.control_policy = f_control_policy(**control_policy_par)
        self.cost_function = f_cost_function(**cost_function_par)
        self.state_samples_history, self.input_samples_history, self.noiseless_states_history = [], [], []
        self.num_data_collection = 0
        self.log_path = log_path
        if log_path is not None:
            self.log_dict = {}
        self._trial_index = None
        self._seed_base = None
        self._rollouts = 0

    # ---- noise bookkeeping ---------------------------------------------------------------------------------------------
    def _next_seed(self):
        """Philox key of the next rollout: a base drawn once from torch's RNG (rank 0's, under torch.distributed) plus a
        rollout counter, so every rank uses the same key and `torch.manual_seed` makes runs reproducible."""
        if self._seed_base is None:
            base = torch.randint(0, 2 ** 60, (1,), dtype=torch.int64)
            rank, world, group = D.world()
            if world > 1:
                b = base.to(self.device)
                torch.distributed.broadcast(b, src=0, group=group)
                base = b.cpu()
            self._seed_base = int(base.item())
        self._rollouts += 1
        return (self._seed_base + GOLDEN * self._rollouts) & MASK64

    def _initial_particles(self, mean, var, flg_uniform, up, low, flg_multi, count, offset, seed, noise, seed_dev=None):
        """Initial particle cloud (reference :635-657): Gaussian, uniform or multi-modal Gaussian."""
        dev = self.device
        as_dev = lambda v: torch.as_tensor(v, dtype=self.dtype, device=dev)  # noqa: E731
        if noise is not None and noise.get("x0") is not None:
            return as_dev(noise["x0"])
        if flg_uniform:
            return ops.init_particles("uniform", as_dev(low).reshape(1, -1), as_dev(up).reshape(1, -1), count, seed, offset, seed_dev)
        mean, std = as_dev(mean), torch.sqrt(as_dev(var))
        if noise is not None and noise.get("eps0") is not None and not flg_multi:
            return mean.reshape(1, -1) + std.reshape(1, -1) * as_dev(noise["eps0"])
        if not flg_multi:
            mean, std = mean.reshape(1, -1), std.reshape(1, -1)
        return ops.init_particles("gauss", mean, std, count, seed, offset, seed_dev)

    def _meas_struct(self):
        return None

    # ---- the particle rollout ------------------------------------------------------------------------------------------
    def apply_policy(self, particles_initial_state_mean, particles_initial_state_var, flg_particles_init_uniform,
                     particles_init_up_bound, particles_init_low_bound, flg_particles_init_multi_gauss, num_particles, T_control,
                     p_dropout=0.0, _noise=None):
        """Simulate `num_particles` particles for `T_control` steps under the current policy (reference :615-674).
        Returns states [H, M_local, Ds] and inputs [H, M_local, Du] (M_local = num_particles unless sharded over ranks).
        `_noise` (tests only) injects {"x0" | "eps0", "eps", "masks", "meas_eps"} instead of the Philox streams."""
        H, Ds, Du = int(T_control), self.state_dim, self.input_dim
        pol, ml = self.control_policy, self.model_learning
        rank, world, group = D.world()
        if int(num_particles) < world:
            raise RuntimeError("apply_policy: %d particles cannot be sharded over %d ranks" % (int(num_particles), world))
        offset, count = D.shard(num_particles, rank, world)
        seed = self._next_seed()
        x0 = self._initial_particles(particles_initial_state_mean, particles_initial_state_var, flg_particles_init_uniform,
                                     particles_init_up_bound, particles_init_low_bound, flg_particles_init_multi_gauss, count, offset,
                                     seed, _noise)
        params = [pol.log_lengthscales, pol.centers, pol.f_linear.weight] + ([pol.f_linear.bias] if pol.flg_bias else [])
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or x0.requires_grad)
        call = self._rollout_call(x0, count, offset, H, p_dropout, seed, _noise, int(num_particles))
        fused = call.fused
        states, inputs, cost, std = _ParticleRollout.apply(call, need_grad, x0, (rank, world, group, int(num_particles)), *params)
        if world > 1:
            states._mcp_shard = (rank, world, group, int(num_particles))  # Expected_cost's generic path merges across ranks with it
        if fused is not None:
            key = self._trial_index if getattr(self.cost_function, "flg_var_lengthscales", False) else None
            states._mcp_fused_cost = (self.cost_function, key, cost, std)
        return states, inputs

    def _rollout_call(self, x0, count, offset, H, p_dropout, seed, noise, num_particles, seed_dev=None):
        """Flatten the current model / policy / cost into the argument bundle of torch.ops.mcpilco.rollout_fwd / rollout_bwd."""
        from .. import torch_ops as TO
        Ds, Du = self.state_dim, self.input_dim
        pol, ml = self.control_policy, self.model_learning
        cf = self.cost_function
        fused = (cf.fused_spec_cached(Ds, H, self._trial_index) if hasattr(cf, "fused_spec_cached") else
                 cf.fused_spec(Ds, H, self._trial_index) if hasattr(cf, "fused_spec") else None)
        cst, ctraj = fused if fused is not None else (None, None)
        if ctraj is not None:
            ctraj = torch.as_tensor(ctraj, dtype=self.dtype, device=self.device)
        nz = noise or {}
        call = TO.RolloutCall(ml.rollout_model_struct(Ds, Du), ml.fitted_gps(), pol.policy_struct(), pol.policy_tensors(), cost=cst,
                              cost_traj=ctraj, meas=self._meas_struct(), M=count, H=H, p_dropout=p_dropout, seed=seed, particle_offset=offset,
                              eps=nz.get("eps"), masks=nz.get("masks"), meas_eps=nz.get("meas_eps"), M_global=int(num_particles),
                              seed_dev=seed_dev)
        call.fused = fused
        return call

    def rollout(self, data_collection_index, T_rollout=None, particle_pred=False):
        """Open-loop model rollout along a recorded input trajectory (reference :347-373): one particle, mean prediction."""
        st = self.state_samples_history[data_collection_index]
        T_rollout = st.shape[0] if T_rollout is None else T_rollout
        x = torch.tensor(st[0:1, :], dtype=self.dtype, device=self.device)
        u = torch.tensor(self.input_samples_history[data_collection_index], dtype=self.dtype, device=self.device)
        traj = torch.zeros([T_rollout, self.state_dim], dtype=self.dtype, device=self.device)
        traj[0:1, :] = x
        with torch.no_grad():
            for t in range(1, T_rollout):
                traj[t:t + 1, :], _, _ = self.model_learning.get_next_state(current_state=traj[t - 1:t, :], current_input=u[t - 1:t, :],
                                                                            particle_pred=particle_pred)
        return traj.cpu().numpy()

    # ---- the optimisation loop -----------------------------------------------------------------------------------------
    def reinforce_policy(self, T_control, num_particles, trial_index, particles_initial_state_mean, particles_initial_state_var,
                         flg_particles_init_uniform, particles_init_up_bound, particles_init_low_bound, flg_particles_init_multi_gauss,
                         opt_steps_list, lr_list, f_optimizer, num_step_print=10, policy_reinit_dict=None, p_dropout_list=None,
                         std_cost_filt_order=None, std_cost_filt_cutoff=None, max_std_cost=None, alpha_cost=0.99, alpha_input=0.99,
                         alpha_diff_cost=0.99, lr_reduction_ratio=0.5, lr_min=0.001, p_drop_reduction=0.0, min_diff_cost=0.1,
                         num_min_diff_cost=200, min_step=np.inf, max_reinit=np.inf):
        """Gradient-based policy improvement on the particle cost (reference :375-613): optimiser built from the eval'd
        `f_optimizer` string, NaN re-sampling (<= 10 attempts) and policy re-initialisation, exponential monitors of the cost
        decrease driving learning-rate halving, dropout reduction and early exit.  `max_reinit` (not in the reference, default
        unlimited like the reference) bounds the number of NaN-triggered policy re-initialisations."""
        H = int(T_control / self.T_sampling)
        n_steps = opt_steps_list[trial_index]
        p_drop0 = 0.0 if p_dropout_list is None else p_dropout_list[trial_index]
        self._trial_index = trial_index
        init = dict(particles_initial_state_mean=particles_initial_state_mean, particles_initial_state_var=particles_initial_state_var,
                    flg_particles_init_uniform=flg_particles_init_uniform, flg_particles_init_multi_gauss=flg_particles_init_multi_gauss,
                    particles_init_up_bound=particles_init_up_bound, particles_init_low_bound=particles_init_low_bound,
                    num_particles=num_particles, T_control=H)
        f_optim = eval(f_optimizer)

        graphed = [None]

        def graph_for(p_drop):
            """The captured fwd + bwd rollout for the current policy tensors / dropout rate, or None (sharded run, user cost, switched off)."""
            if not _GraphedRollout.eligible(self):
                return None
            if graphed[0] is None or graphed[0].signature != _GraphedRollout.signature_of(self, init, p_drop):
                graphed[0] = _GraphedRollout(self, init, p_drop)
            return graphed[0]

        def sample(p_drop):
            """Rollout + cost, re-sampled up to 10 times while the cost is NaN; returns (states, inputs, cost, std, still_nan, graph).
            With a captured graph the rollout's backward pass has already run when this returns (gradients wait in the graph's
            buffers); without one the caller runs cost.backward()."""
            g = graph_for(p_drop)
            for _ in range(10):
                if g is not None:
                    cost, std = g.replay()
                    states, inputs = g.states, g.inputs
                else:
                    states, inputs = self.apply_policy(p_dropout=p_drop, **init)
                    cost, std = self.cost_function(states, inputs, trial_index)
                if not bool(torch.isnan(cost)):  # the one host synchronisation per step, as in the reference (MC_PILCO.py:497)
                    return states, inputs, cost, std, False, g
            return states, inputs, cost, std, True, g

        def fresh():
            z = lambda n: torch.zeros(n, device=self.device, dtype=self.dtype)  # noqa: E731
            return dict(cost=z(n_steps), std=z(n_steps), es1=z(n_steps + 1), es2=0.0, ratio=z(n_steps + 1), lr=lr_list[trial_index],
                        p_drop=p_drop0, min_diff=min_diff_cost, min_step=min_step, prev=0.0, step=0, done=0)

        with torch.no_grad():  # cost level before optimisation: initialises the cost-difference filter
            for _ in range(10):
                st0, in0 = self.apply_policy(p_dropout=p_drop0, **init)
                cost0, _ = self.cost_function(st0, in0, trial_index)
                if not bool(torch.isnan(cost0)):
                    break
                print("\nSE filter initialization: Cost is NaN - reinit the policy")
                self.control_policy.reinit(**policy_reinit_dict)
            del st0, in0
        s = fresh()
        cost_tm1 = cost0
        optimizer = f_optim(p=self.control_policy.parameters(), lr=s["lr"])
        reinit_counter, t_start = 0, time.time()
        a = alpha_diff_cost
        while s["step"] < n_steps:
            optimizer.zero_grad()
            states, inputs, cost, std, is_nan, g = sample(s["p_drop"])
            k = s["step"]
            s["cost"][k], s["std"][k] = cost.detach(), std.detach()
            with torch.no_grad():
                dc = cost - cost_tm1
                s["es1"][k + 1] = a * s["es1"][k] + (1 - a) * dc
                s["es2"] = a * (s["es2"] + (1 - a) * (dc - s["es1"][k]) ** 2)
                cost_tm1 = s["cost"][k]
                s["ratio"][k + 1] = a * s["ratio"][k] + (1 - a) * (s["es1"][k + 1] / s["es2"].sqrt())
            if g is not None:
                g.deposit_grads()  # the captured backward pass already ran
            else:
                cost.backward()
            optimizer.step()
            if k % num_step_print == 0:
                c = float(cost.detach())
                print("\nOptimization step:", k, "| cost:", c, "| improvement:", s["prev"] - c, "| p_dropout:", s["p_drop"],
                      "| diff_cost_ratio:", float(torch.abs(s["ratio"][k + 1])), "| time:", time.time() - t_start)
                s["prev"], t_start = c, time.time()
            if k > s["min_step"]:
                window = torch.abs(s["ratio"][k + 1 - num_min_diff_cost:k + 1]) < s["min_diff"]
                if int(window.sum()) >= num_min_diff_cost:
                    if s["lr"] > lr_min:
                        s["lr"] = max(s["lr"] * lr_reduction_ratio, lr_min)
                        s["min_diff"] = max(s["min_diff"] / 2, 0.01)
                        s["min_step"] = k + num_min_diff_cost
                        optimizer = f_optim(p=self.control_policy.parameters(), lr=s["lr"])
                        s["p_drop"] = max(s["p_drop"] - p_drop_reduction, 0.0)
                        print("\nstep", k, ": learning rate ->", s["lr"], ", p_dropout ->", s["p_drop"])
                    else:
                        print("\nEXIT FROM OPTIMIZATION: diff_cost_ratio < min_diff_cost for num_min_diff_cost steps")
                        s["step"] = n_steps
            s["step"] += 1
            s["done"] += 1
            if is_nan:  # ten NaN rollouts in a row: new random policy, restart the optimisation
                reinit_counter += 1
                if reinit_counter > max_reinit:
                    raise RuntimeError("reinforce_policy: the particle cost stayed NaN through %d policy re-initialisations" % (reinit_counter - 1))
                print("\nCost is NaN: re-initialize control policy [attempt #" + str(reinit_counter) + "]")
                self.control_policy.reinit(**policy_reinit_dict)
                s = fresh()
                optimizer = f_optim(p=self.control_policy.parameters(), lr=s["lr"])
        n = s["done"]
        return (s["cost"][:n].cpu().numpy(), s["std"][:n].cpu().numpy(), states.detach().cpu().numpy(), inputs.detach().cpu().numpy())

    # ---- host orchestration that stays with the reference --------------------------------------------------------------
    def _out_of_scope(self, *a, **k):
        raise NotImplementedError("trial loop / data collection / logging are host orchestration outside the rollout hot path; the "
                                  "reference's own MC_PILCO keeps them and binds to apply_policy / reinforce_policy (INTEGRATION.md)")

    reinforce = get_data_from_system = get_model_learning_performance = get_rollout_prediction_performance = _out_of_scope
    load_policy_from_log = load_model_from_log = _out_of_scope


class MC_PILCO4PMS(MC_PILCO):
    """MC-PILCO for partially measurable systems (reference :754-962): inside the rollout the policy sees a simulated
    measurement — noisy positions, finite-difference velocities, first-order Butterworth low-pass — while the cost is taken on
    the true states.  Differences from the reference: `std_meas_noise_sim` is honoured when given (the reference only assigns the
    attribute when the argument is None, :805-806, and would raise AttributeError otherwise)."""

    def __init__(self, T_sampling, state_dim, input_dim, f_sim, f_model_learning, model_learning_par, f_rand_exploration_policy,
                 rand_exploration_policy_par, f_control_policy, control_policy_par, f_cost_function, cost_function_par, pos_indeces,
                 vel_indeces, std_meas_noise=None, log_path=None, filtering_dict={}, std_meas_noise_sim=None, dtype=torch.float64,
                 device=torch.device("cuda")):
        super().__init__(T_sampling=T_sampling, state_dim=state_dim, input_dim=input_dim, f_sim=f_sim, f_model_learning=f_model_learning,
                         model_learning_par=model_learning_par, f_rand_exploration_policy=f_rand_exploration_policy,
                         rand_exploration_policy_par=rand_exploration_policy_par, f_control_policy=f_control_policy,
                         control_policy_par=control_policy_par, f_cost_function=f_cost_function, cost_function_par=cost_function_par,
                         std_meas_noise=std_meas_noise, log_path=log_path, dtype=dtype, device=device)
        self.filtering_dict = filtering_dict
        self.pos_indeces, self.vel_indeces = pos_indeces, vel_indeces
        self.std_meas_noise_sim = self.std_meas_noise if std_meas_noise_sim is None else std_meas_noise_sim

    def _meas_struct(self):
        std_pos = np.asarray(self.std_meas_noise_sim, dtype=np.float64)[list(self.pos_indeces)]
        return P.meas_struct(self.pos_indeces, self.vel_indeces, std_pos, self.filtering_dict["fc"], self.T_sampling)

    def apply_policy(self, particles_initial_state_mean, particles_initial_state_var, flg_particles_init_uniform,
                     particles_init_up_bound, particles_init_low_bound, flg_particles_init_multi_gauss, num_particles, T_control,
                     p_dropout=0.0, _noise=None):
        """Rollout with the measurement model in the loop (reference :808-906)."""
        return super().apply_policy(particles_initial_state_mean, particles_initial_state_var, flg_particles_init_uniform,
                                    particles_init_up_bound, particles_init_low_bound, flg_particles_init_multi_gauss, num_particles,
                                    T_control, p_dropout=p_dropout, _noise=_noise)

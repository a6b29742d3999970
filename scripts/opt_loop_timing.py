"""Wall time per optimisation step of MC_PILCO.reinforce_policy at the real cart-pole shape (C1: N=300, M=400, H=60)."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from mcpilco_b200 import workloads as W
dev = torch.device("cuda:0")
sc = W.cartpole_sweep(300, sigma_n=float(np.exp(-4.2)))
obj = bench.build_objects(sc, dev)
with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    for e in range(2):
        obj.model_learning.pretrain_gp(e)
obj.model_learning.set_eval_mode()
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)
kw = dict(T_control=3.0, num_particles=400, trial_index=0, particles_initial_state_mean=T(sc["x0_mean"]), particles_initial_state_var=T(sc["x0_var"]),
          flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None, flg_particles_init_multi_gauss=False,
          lr_list=[0.01], f_optimizer="lambda p, lr : torch.optim.Adam(p, lr)", num_step_print=10 ** 6, p_dropout_list=[0.25],
          policy_reinit_dict=dict(lenghtscales_par=np.ones(5), centers_par=np.ones((200, 5)), weight_par=1.0))
for steps in (20, 200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = obj.reinforce_policy(opt_steps_list=[steps], **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("reinforce_policy %d steps: %.2f ms per optimisation step (cost %.3f -> %.3f)" % (steps, 1e3 * dt / (steps + 1), out[0][0], out[0][-1]), flush=True)

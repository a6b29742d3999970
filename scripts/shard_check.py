"""torchrun --nproc-per-node G scripts/shard_check.py : 1-vs-G GPU consistency.
Every rank runs the SAME global rollout sharded over G ranks through MC_PILCO.apply_policy / cost / backward; rank 0 also
runs it unsharded on its own GPU and compares cost, std and policy gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-pilco_b200")); sys.path.insert(0, ROOT)
import contextlib
import numpy as np, torch, torch.distributed as dist
import bench
from mcpilco_b200 import distributed as D, workloads as W

def run(obj, sc, M, H, dev, seed):
    pol = obj.control_policy
    for p in pol.parameters():
        p.grad = None
    obj._seed_base, obj._rollouts = seed, 0
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)
    st, inp = obj.apply_policy(particles_initial_state_mean=T(sc["x0_mean"]), particles_initial_state_var=T(sc["x0_var"]),
                               flg_particles_init_uniform=False, particles_init_up_bound=None, particles_init_low_bound=None,
                               flg_particles_init_multi_gauss=False, num_particles=M, T_control=H, p_dropout=sc["p_dropout"])
    cost, std = obj.cost_function(st, inp, 0)
    cost.backward()
    return float(cost.detach()), float(std), [p.grad.clone() for p in (pol.log_lengthscales, pol.centers, pol.f_linear.weight)], st

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc = W.cartpole_sweep(512)
    obj = bench.build_objects(sc, dev)
    with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
        for e in range(sc["E"]):
            obj.model_learning.pretrain_gp(e)
    obj.model_learning.set_eval_mode()
    M, H = 4099, 12          # not divisible by the world size on purpose
    c_sh, s_sh, g_sh, st_sh = run(obj, sc, M, H, dev, seed=1234)
    off, cnt = D.shard(M, rank, world)
    ok = True
    if rank == 0:
        import mcpilco_b200.policy_learning.MC_PILCO as MCP
        saved = D.world
        D.world = lambda: (0, 1, None)                      # same process, unsharded
        try:
            c1, s1, g1, st1 = run(obj, sc, M, H, dev, seed=1234)
        finally:
            D.world = saved
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        errs = dict(cost=abs(c_sh - c1) / abs(c1), std=abs(s_sh - s1) / abs(s1), g_ls=rel(g_sh[0], g1[0]), g_c=rel(g_sh[1], g1[1]), g_w=rel(g_sh[2], g1[2]),
                    states_bitwise=bool(torch.equal(st_sh, st1[:, off:off + cnt])))
        ok = errs["cost"] < 1e-12 and errs["std"] < 1e-10 and max(errs["g_ls"], errs["g_c"], errs["g_w"]) < 1e-10 and errs["states_bitwise"]
        print("shard_check world=%d M=%d H=%d: %s -> %s" % (world, M, H, errs, "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

main()

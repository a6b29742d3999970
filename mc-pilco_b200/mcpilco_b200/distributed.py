"""Particle data-parallelism: one process per GPU, GP pack and policy replicated, particles sharded.

The only collectives of the path: an all-gather of the per-step cost statistics [H, 2] (mean, M2) and a SUM
all-reduce of the flat policy gradient.  Both are tiny (<= 100 KB); NCCL over NVLink on the GPU box, gloo in the CPU tests
of this host logic.  Philox counters are keyed by the global particle id, so a sharded rollout draws exactly the noise the
single-GPU rollout draws.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size, group) of the default process group, or (0, 1, None) when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), dist.group.WORLD
    return 0, 1, None


def shard(num_particles, rank, world_size):
    """Contiguous particle range [offset, offset + count) of `rank`; the first (num_particles % world_size) ranks get one more."""
    base, rem = divmod(int(num_particles), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def merge_cost_stats(stats, counts):
    """Merge per-rank per-step (mean, M2 = sum (c - mean)^2) into global (mean, M2) with Chan's pairwise update, in rank
    order (deterministic).  stats: [G, H, 2] tensor, counts: list of G particle counts."""
    n = float(counts[0])
    mean, m2 = stats[0, :, 0].clone(), stats[0, :, 1].clone()
    for g in range(1, stats.shape[0]):
        nb = float(counts[g])
        if nb == 0:
            continue
        delta = stats[g, :, 0] - mean
        tot = n + nb
        mean = mean + delta * (nb / tot)
        m2 = m2 + stats[g, :, 1] + delta * delta * (n * nb / tot)
        n = tot
    return mean, m2


def expected_cost_from_stats(mean, m2, num_particles):
    """sum_t mean_t and sum_t sqrt(M2_t / (M - 1)): Expected_cost.forward (reference Cost_function.py:33-36)."""
    return mean.sum(), torch.sqrt(m2 / (num_particles - 1)).sum()


def gather_cost_stats(local_stats, group, world_size):
    out = [torch.empty_like(local_stats) for _ in range(world_size)]
    dist.all_gather(out, local_stats.contiguous(), group=group)
    return torch.stack(out)


def allreduce_sum_(flat, group):
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat

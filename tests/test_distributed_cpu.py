"""CPU (gloo, world_size 2): the host logic of particle sharding — shard arithmetic, the merge of per-rank cost statistics
and the two collectives of the path (all-gather of [H,2] stats, SUM all-reduce of the flat gradient)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcpilco_b200 import distributed as D


def test_shard_partitions_exactly():
    for M in (1, 7, 400, 8192, 1000003):
        for G in (1, 2, 3, 8):
            parts = [D.shard(M, r, G) for r in range(G)]
            assert sum(c for _, c in parts) == M
            assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(G - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_merge_cost_stats_matches_global_moments():
    rs = np.random.RandomState(0)
    H, counts = 6, [5, 3, 9, 1]
    costs = [rs.rand(H, c) for c in counts]
    stats = torch.tensor(np.stack([np.stack([c.mean(1), ((c - c.mean(1, keepdims=True)) ** 2).sum(1)], 1) for c in costs]))
    mean, m2 = D.merge_cost_stats(stats, counts)
    allc = np.concatenate(costs, 1)
    np.testing.assert_allclose(mean.numpy(), allc.mean(1), rtol=1e-14)
    np.testing.assert_allclose(m2.numpy(), ((allc - allc.mean(1, keepdims=True)) ** 2).sum(1), rtol=1e-13)
    cost, std = D.expected_cost_from_stats(mean, m2, allc.shape[1])
    t = torch.tensor(allc)
    np.testing.assert_allclose(float(cost), float(t.mean(1).sum()), rtol=1e-14)
    np.testing.assert_allclose(float(std), float(torch.std(t, 1).sum()), rtol=1e-13)   # unbiased, like Expected_cost


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, M, H, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, group = D.world()
        assert (r, w) == (rank, world)
        rs = np.random.RandomState(1)
        costs = torch.tensor(rs.rand(H, M))                     # the global per-particle costs, identical on every rank
        grads = torch.tensor(rs.randn(M, 11))                   # per-particle gradient contributions
        off, cnt = D.shard(M, rank, world)
        c = costs[:, off:off + cnt]
        local = torch.stack([c.mean(1), ((c - c.mean(1, keepdim=True)) ** 2).sum(1)], 1)
        stats = D.gather_cost_stats(local, group, world)
        mean, m2 = D.merge_cost_stats(stats, [D.shard(M, q, world)[1] for q in range(world)])
        cost, std = D.expected_cost_from_stats(mean, m2, M)
        flat = grads[off:off + cnt].sum(0) / M
        D.allreduce_sum_(flat, group)
        ok = (abs(float(cost) - float(costs.mean(1).sum())) < 1e-12 and abs(float(std) - float(torch.std(costs, 1).sum())) < 1e-12
              and float((flat - grads.sum(0) / M).abs().max()) < 1e-12)
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_collectives():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_world():
    assert D.world() == (0, 1, None)

"""World-size-2 gloo test of the data-parallel host logic (runs on CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_frame_super_resolution_b200 import dp


def test_shard_bursts_partition():
    for n, g in ((256, 8), (10, 4), (3, 8), (0, 2)):
        got = sorted(b for r in range(g) for b in dp.shard_bursts(n, r, g))
        assert got == list(range(n))
        sizes = [len(dp.shard_bursts(n, r, g)) for r in range(g)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_bursts(4, 2, 2)
    assert dp.burst_seed(1234, 7) == 1241


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = dp.shard_bursts(5, rank, world)
    # rank r "processes" its bursts: 48.77 MP each, rank 1 is slower
    units, ms, ups = dp.aggregate_throughput(48.77 * len(mine), 100.0 * (rank + 1))
    seeds = [dp.burst_seed(1234, b) for b in mine]
    q.put((rank, mine, seeds, units, ms, ups))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, s0, u0, ms0, ups0), (r1, b1, s1, u1, ms1, ups1) = res
    assert b0 == [0, 2, 4] and b1 == [1, 3] and s0 == [1234, 1236, 1238]
    assert u0 == u1 == pytest.approx(48.77 * 5) and ms0 == ms1 == 200.0      # SUM of units, MAX of time
    assert ups0 == pytest.approx(48.77 * 5 / 0.2)


def test_aggregate_without_group():
    u, ms, ups = dp.aggregate_throughput(10.0, 50.0)
    assert (u, ms) == (10.0, 50.0) and ups == pytest.approx(200.0)

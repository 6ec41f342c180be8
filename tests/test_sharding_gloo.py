"""N>1 host logic on CPU: two gloo ranks shard the global environment range without gaps or overlap
and combine episode statistics with exactly one all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import import_package


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_envs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = import_package()
    start, stop = pkg.shard_range(global_envs, rank, world)
    # every rank reports the ids it owns; rank 0 checks the partition
    owned = torch.zeros(global_envs, dtype=torch.int32)
    owned[start:stop] = 1
    dist.all_reduce(owned)
    # fake per-device episode sums: rank r finished (r+1)*10 episodes with return 100*(r+1) each
    n = (rank + 1) * 10
    sums = torch.zeros(21, dtype=torch.float64)
    sums[0] = 100.0 * (rank + 1) * n      # return
    sums[1] = 1500.0 * n                  # length
    sums[6] = 50.0 * n                    # ndr
    sums[20] = n
    calls = {"n": 0}
    real = dist.all_reduce

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    dist.all_reduce = counting
    try:
        stats = pkg.allreduce_episode_stats(sums)
    finally:
        dist.all_reduce = real
    if rank == 0:
        q.put((owned.tolist(), stats, calls["n"], (start, stop)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("global_envs", [4096, 4097])
def test_two_ranks_partition_and_single_allreduce(global_envs):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    owned, stats, ncalls, r0 = q.get()
    assert owned == [1] * global_envs                       # disjoint and complete
    assert r0 == (0, 2 * (((global_envs + 1) // 2 + 1) // 2))   # whole environment pairs: rank 0 owns ceil(pairs / 2) of them
    assert ncalls == 1                                       # the path's only collective
    assert stats["episodes"] == 30.0
    assert abs(stats["return"] - (100.0 * 10 + 200.0 * 20) / 30.0) < 1e-12
    assert stats["length"] == 1500.0 and stats["ndr"] == 50.0


def test_shard_range_properties():
    pkg = import_package()
    for E in (16, 1000, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            spans = [pkg.shard_range(E, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == E
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 3
            assert all(a % 2 == 0 for a, _ in spans)        # environment pairs never straddle shards
    with pytest.raises(ValueError):
        pkg.shard_range(4, 4, 4)
    with pytest.raises(ValueError):
        pkg.shard_range(3, 0, 8)
    with pytest.raises(ValueError):
        pkg.shard_range(5, 0, 4)                            # 3 pairs cannot feed 4 ranks


def test_single_process_stats_need_no_process_group():
    pkg = import_package()
    sums = torch.zeros(21, dtype=torch.float64)
    sums[0], sums[20] = 500.0, 5.0
    out = pkg.allreduce_episode_stats(sums)
    assert out["episodes"] == 5.0 and out["return"] == 100.0
    assert pkg.allreduce_episode_stats(torch.zeros(21, dtype=torch.float64))["return"] == 0.0
    with pytest.raises(ValueError):
        pkg.allreduce_episode_stats(torch.zeros(5))

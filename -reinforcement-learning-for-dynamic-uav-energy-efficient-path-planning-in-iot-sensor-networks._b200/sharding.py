"""Multi-GPU sharding of the batched environments.

Environments never interact (the reference's DummyVecEnv just loops over them, dqn.py:1276), so the
global id range [0, E) is cut into contiguous near-equal shards (whole environment pairs), one per rank / GPU; the step needs no
communication.  The only exchange is one all-reduce(sum) of the episode-statistic sums that the step
kernel accumulates on each device (`UAVVecEnv.stat_sums`, float64[21]) -- the batched analogue of
`CurriculumCallback._on_step` averaging `last_episode_stats` over workers (dqn.py:929-1039).  It is a
latency-only collective (168 bytes) issued per reporting interval, off the step's critical path.

All functions work with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo on CPU in
the tests (world_size 2).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ._native import STAT_COUNT, STAT_NAMES, STAT_SUMS


def shard_range(global_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of global env ids owned by `rank`.  Shards are cut in units of environment PAIRS
    (ids 2m, 2m+1): with up to 32 sensors per environment the observation draws of a pair come from one Philox
    call, so a pair must live on one device and every shard starts at an even id (`env_id_offset` even).  Shard
    sizes differ by at most 3 (one pair, plus the unpaired last environment of an odd total)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    if global_envs < world:
        raise ValueError("fewer environments (%d) than ranks (%d)" % (global_envs, world))
    pairs = (global_envs + 1) // 2
    if pairs < world:
        raise ValueError("fewer environment pairs (%d) than ranks (%d)" % (pairs, world))
    base, extra = divmod(pairs, world)
    first = rank * base + min(rank, extra)
    last = first + base + (1 if rank < extra else 0)
    return 2 * first, min(2 * last, global_envs)


def make_sharded_env(pkg_env_cls, global_envs: int, rank: Optional[int] = None, world: Optional[int] = None,
                     device: Optional[str] = None, **env_kwargs):
    """Construct this rank's UAVVecEnv: same seed on every rank, `env_id_offset` = first owned id, so a
    given global environment draws the same Philox stream whatever the number of GPUs."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    start, stop = shard_range(global_envs, rank, world)
    if device is None:
        device = "cuda:%d" % (rank % max(torch.cuda.device_count(), 1))
    return pkg_env_cls(stop - start, env_id_offset=start, device=device, **env_kwargs)


def allreduce_episode_stats(stat_sums: torch.Tensor, group=None) -> Dict[str, float]:
    """Sum the per-device episode-statistic vectors over all ranks (ONE collective) and return the
    job-wide means per finished episode.  `stat_sums` is float64[21]: sums of the 20 statistics and,
    last, the number of finished episodes.  Returns {"episodes": n, "<stat>": mean, ...}."""
    if stat_sums.numel() != STAT_SUMS:
        raise ValueError("expected %d sums, got %d" % (STAT_SUMS, stat_sums.numel()))
    total = stat_sums.detach().clone().to(torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    total = total.cpu()
    n = float(total[STAT_COUNT])
    out = {"episodes": n}
    for i, name in enumerate(STAT_NAMES):
        out[name] = float(total[i]) / n if n > 0 else 0.0
    return out

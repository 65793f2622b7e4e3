"""Multi-GPU layout of the path: environments are independent, so they shard
as contiguous blocks of the GLOBAL env index with no per-step communication; seeds stay
base_seed + global index, so results do not depend on the number of ranks.  The only collective is the
end-of-rollout all-reduce of the int64[8] episode counters (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous [lo, hi) of global env indices owned by `rank` (the first n_total % world ranks get one more)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_episode_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the int64[8] counters of DeliveryBatch.episode_stats() over all ranks (exact: integers)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def make_sharded_env(n_total: int, rank: int = None, world: int = None, **kwargs):
    """VectorizedEnv holding this rank's shard of an n_total-env job."""
    from .env import Environment
    from .env_vectorized import VectorizedEnv

    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(n_total, rank, world)
    return VectorizedEnv(Environment, hi - lo, first_env_index=lo, **kwargs)

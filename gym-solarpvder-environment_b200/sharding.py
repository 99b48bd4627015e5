"""Multi-GPU layout: environments are sharded contiguously by global index, one process per GPU.

There is no collective on the hot path (envs are independent DERs, reference README.md:6).  RNG
streams are keyed by the GLOBAL env index (env_offset + local index), so env i produces the same
events and trajectory whichever rank owns it.  The only communication is the optional reduction
of the 16-double episode-statistics vector (reference gym_PVDER/envs/env_utilities.py:12-46
counters) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

STATS_FIELDS = ["return_sum", "steps_sum", "n_done", "n_failed", "act0", "act1", "act2", "act3", "act4",
                "windup_sub_steps", "n_envs", "exact_sub_steps"]


def shard_bounds(total_envs: int, rank: int, world_size: int):
    """[lo, hi) of the envs owned by ``rank``: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_sharded_env(total_envs, rank=None, world_size=None, device=None, **kwargs):
    """PVDERVecEnv over this rank's shard (rank/world default to torch.distributed's)."""
    import torch
    import torch.distributed as dist

    from .envs.vec_env import PVDERVecEnv

    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(total_envs, rank, world_size)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return PVDERVecEnv(hi - lo, device=device, env_offset=lo, **kwargs)


def reduce_stats(stats, group=None):
    """Sum the per-shard statistics vector over all ranks (in place) and return it as a dict."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    vals = stats.detach().cpu().tolist()
    return dict(zip(STATS_FIELDS, vals))

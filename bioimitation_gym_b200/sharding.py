"""Multi-GPU layout of the path: envs are independent, so the batch is sharded
by global env index over one process per GPU and NOTHING crosses NVLink in
``reset``/``step``.  The reference's only parallel axis is the same one (RLlib
rollout workers = one OpenSim model per process, reference
``configs/train_default.py:22``).  The single collective is an all-gather of
the 16-double rollout-statistics vector (``bio_stats``) per reporting interval,
NCCL on GPUs / gloo on CPU, outside the step path.
"""
from __future__ import annotations

import os
from typing import Any, Mapping, Optional, Tuple


def shard(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(local count, global offset) of a contiguous, balanced split."""
    base, rem = divmod(int(n_total), int(world))
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return cnt, off


def env_from_torchrun():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def all_gather_stats(stats):
    """[16] float64 tensor per rank -> [world, 16] on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats.reshape(1, -1).clone()
    out = [torch.zeros_like(stats) for _ in range(dist.get_world_size())]
    dist.all_gather(out, stats)
    return torch.stack(out)


def make_sharded_vec_env(env_id: str, n_total: int, config: Optional[Mapping[str, Any]] = None):
    """This rank's slice of an n_total-env batch (launch one process per GPU with
    torchrun); results are identical to the single-GPU batch because the reset RNG
    is keyed by the global env index (``env_offset``)."""
    from .backend import VecEnv
    rank, world, local_rank = env_from_torchrun()
    cnt, off = shard(n_total, rank, world)
    cfg = dict(config or {})
    cfg.update(num_envs=cnt, env_offset=off, device=local_rank)
    return VecEnv(env_id, cfg)

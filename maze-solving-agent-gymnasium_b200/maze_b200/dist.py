"""Multi-GPU plumbing: one process per GPU, envs and maze slots sharded by index, no per-step
collective.  The only exchange is the end-of-rollout reduction of episode statistics
(torch.distributed: NCCL on GPUs, gloo in the CPU tests of this host logic)."""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

STAT_NAMES = ("episodes", "wins", "truncations", "steps", "return_sum")


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    start: int    # first global index owned by this rank
    count: int    # number of indices owned

    @property
    def stop(self):
        return self.start + self.count


def shard_range(total: int, rank: int, world: int) -> Shard:
    """Contiguous partition of [0, total): rank r owns [r*total/world, (r+1)*total/world);
    the first total % world ranks get one extra index."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    start = rank * base + min(rank, extra)
    return Shard(rank, world, start, base + (1 if rank < extra else 0))


def env_from_torchrun():
    """(rank, local_rank, world) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def reduce_statistics(local: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a float64 [len(STAT_NAMES)] statistics vector over all ranks (identity when
    torch.distributed is not initialised)."""
    if local.dtype != torch.float64 or local.numel() != len(STAT_NAMES):
        raise ValueError("statistics vector must be float64 of length %d" % len(STAT_NAMES))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        local = local.clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def statistics_dict(vec: torch.Tensor) -> dict:
    v = vec.cpu().tolist()
    out = {k: (int(round(x)) if k != "return_sum" else float(x)) for k, x in zip(STAT_NAMES, v)}
    out["win_rate"] = out["wins"] / out["episodes"] if out["episodes"] else float("nan")
    return out

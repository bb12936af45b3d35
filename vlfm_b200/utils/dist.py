"""Multi-GPU plumbing for the env-sharded step path.

The path shards over independent environments: env ``e`` lives on rank ``e // per_rank``;
there is NO collective on the step path.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used only for the barrier around the timed region, the max-over-ranks of the
elapsed time and the optional all-gather of a small per-rank metrics vector.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def env_shard(rank: int, world_size: int, envs_per_rank: int) -> List[int]:
    """Global environment ids owned by ``rank`` (contiguous blocks, weak scaling)."""
    return list(range(rank * envs_per_rank, (rank + 1) * envs_per_rank))


def owner_of(env_id: int, envs_per_rank: int) -> int:
    return env_id // envs_per_rank


def max_over_ranks(value: float, device: torch.device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_metrics(values: Sequence[float], device: torch.device) -> List[List[float]]:
    """all_gather of a fixed-size per-rank metrics vector (steps, ms, checksums...)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t.tolist()]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def aggregate_throughput(world_size: int, envs_per_rank: int, steps: int, max_ms: float) -> float:
    """Whole-job env-steps per second given the slowest rank's elapsed time."""
    return world_size * envs_per_rank * steps / (max_ms * 1e-3)

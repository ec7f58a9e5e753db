"""Scene sharding across the GPUs of one node.

Every op on the path is per-scene (the reference selects the batch element with blockIdx: sampling_gpu.cu:105,
ball_query_gpu.cu:15, group_points_gpu.cu:59) and eval-mode BN has no cross-sample coupling, so a batch is cut into
contiguous scene ranges, one per rank, weights replicated, and NO collective runs on the data path.  The only
communication is the reporting reduction below (backend NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_scenes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `n_scenes` over `world` ranks; the first n_scenes % world ranks take one extra scene."""
    if world < 1 or not (0 <= rank < world) or n_scenes < 0:
        raise ValueError(f"bad shard request: n_scenes={n_scenes} rank={rank} world={world}")
    base, extra = divmod(n_scenes, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_report(scenes_local: int, ms_local: float, device: torch.device | str = "cpu") -> Tuple[int, float]:
    """(total scenes over all ranks, MAX over ranks of the device-timed milliseconds): the two numbers the whole-job
    throughput is computed from.  One all_reduce(SUM) + one all_reduce(MAX) of 8 bytes each; identity without a group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(scenes_local), float(ms_local)
    n = torch.tensor([float(scenes_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(n.item())), float(t.item())


def throughput(scenes_total: int, ms_max: float) -> float:
    """Whole-job scenes/s."""
    return scenes_total / (ms_max / 1e3) if ms_max > 0 else 0.0

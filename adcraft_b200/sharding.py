"""Env sharding across ranks: contiguous env ranges, global env ids key the Philox counters."""
from __future__ import annotations

import os
from typing import Tuple


def env_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the envs owned by `rank` (remainder spread over the first ranks)."""
    assert 0 <= rank < world_size
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_distributed(backend: str = "nccl"):
    """torch.distributed bootstrap for one process per GPU (torchrun env vars); returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world


def gather_observations(flat, dst: int = 0, group=None, max_bytes: int = 64 << 30):
    """Optional observation gather to the learner GPU (BASELINE.json north_star): every rank's flat
    observation rows ``[E_rank, 5K+2]`` (``VectorBiddingSimulation.flat_observation()``, written by the
    step kernels) are gathered on rank ``dst`` in rank order -- which is global env order, since rank g
    owns envs ``[g * E, (g + 1) * E)`` -- with ONE ``torch.distributed.gather`` (NCCL over NVLink for
    CUDA tensors, gloo for CPU tensors).  Returns ``[world * E_rank, 5K+2]`` on ``dst``, None
    elsewhere.  Refuses what cannot fit one GPU (C5: 200 GB per step; a data-parallel learner is the
    answer there): ``max_bytes`` bounds the gathered tensor."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = world * flat.numel() * flat.element_size()
    if total > max_bytes:
        raise ValueError(f"gather_observations: {total / 1e9:.1f} GB of observations do not fit the learner GPU "
                         f"(limit {max_bytes / 1e9:.1f} GB); keep the learner data-parallel")
    flat = flat.contiguous()
    parts = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
    dist.gather(flat, parts, dst=dst, group=group)
    return torch.cat(parts, dim=0) if rank == dst else None

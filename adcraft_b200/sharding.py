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

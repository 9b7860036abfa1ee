"""Multi-GPU plumbing for the mode the hot path shards naturally in (mode A): independent meshes,
one self-prior fit each, one process per GPU, **no data-path collective**.  torch.distributed (NCCL on the GPU box,
gloo in the CPU tests) is used only for the barrier, the max-over-ranks timing and the final gather of results.
"""
from __future__ import annotations

import os
from typing import Any, List

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_items(n_items: int, rank: int, world: int) -> List[int]:
    """static round-robin of mesh ids over ranks: item i is fitted by rank i % world"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def init(backend: str | None = None, device: torch.device | None = None) -> bool:
    """initialise the default process group from the torchrun environment; returns False when world_size == 1"""
    _, _, world = env_rank_world()
    if world <= 1:
        return False
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return True


def barrier(device: torch.device | None = None) -> None:
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_initialized():
        dist.barrier()
        if device is not None and device.type == "cuda":
            torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    """multi-GPU timings are reported as the max over ranks (device-side events, never wall clock)"""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(obj: Any) -> List[Any]:
    """rank-ordered list of every rank's (picklable) result; on a single process just [obj]"""
    if not dist.is_initialized():
        return [obj]
    out: List[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out

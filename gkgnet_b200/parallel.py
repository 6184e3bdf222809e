"""Data-parallel plumbing for the graph hot path (one process per GPU).

The reference trains with mmcv's ``MMDistributedDataParallel`` over NCCL
(mmcls/apis/train.py:121-125, ``broadcast_buffers=False``) and nothing else: every (image, channel
group) kNN problem is independent, so the graph path itself needs no collective (SURVEY.md
section 8(e)).  Sharding is by image; the only exchange is the gradient all-reduce that DDP runs
behind backward.  These helpers keep that contract in one place so that ``bench.py`` (NCCL, GPUs)
and the CPU tests (gloo, world size 2) exercise the same code.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns ``(world, rank, local_rank)``; a single process is world 1 and creates no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("GKG_PIN_CORES", "0") == "1":
        pin_host_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return world, rank, local


def pin_host_cores(local: int, local_world: int) -> list[int]:
    """Give every rank of the box its own contiguous slice of the host cores this process may use (the launch
    threads of the ranks otherwise migrate over each other).  Opt-in: GKG_PIN_CORES=1."""
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return []
    per = len(cores) // max(local_world, 1)
    if per < 1:
        return cores
    mine = cores[local * per:(local + 1) * per]
    os.sched_setaffinity(0, mine)
    return mine


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced ``[start, stop)`` slice of ``total`` images for ``rank`` (the first
    ``total % world`` ranks take one extra image; every image is owned by exactly one rank)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def barrier(device: torch.device | None = None) -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    """Slowest rank's value (multi-GPU timings are the max over ranks, never a wall clock)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | None = None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def data_parallel(module: torch.nn.Module, device: torch.device | None = None, **kw) -> torch.nn.Module:
    """Wrap ``module`` for synchronous data-parallel training the way the reference does:
    gradient all-reduce only, buffers not broadcast (apis/train.py:124).  ``relative_pos`` tables
    are frozen parameters (requires_grad=False) and take no part in the reduction."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return module
    from torch.nn.parallel import DistributedDataParallel as DDP
    kw.setdefault("broadcast_buffers", False)
    kw.setdefault("gradient_as_bucket_view", True)      # gradients live in the all-reduce buckets: no copy per step
    if device is not None and device.type == "cuda":
        return DDP(module, device_ids=[device.index], output_device=device.index, **kw)
    return DDP(module, **kw)

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


class FlatGradients:
    """The gradients of ``params`` as views of ONE flat fp32 buffer: zeroed with one memset, averaged over the ranks
    with a handful of large all-reduces (what DDP's buckets do: apis/train.py:121-125 semantics, gradient mean over the
    data-parallel ranks), and with stable addresses -- the precondition for capturing the step in a CUDA graph."""

    def __init__(self, params, bucket_bytes: int = 25 * 1024 * 1024):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("parameters must share device and dtype")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.bucket_elems = max(1, bucket_bytes // self.flat.element_size())

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        for chunk in self.flat.split(self.bucket_elems):
            dist.all_reduce(chunk)
        self.flat.mul_(1.0 / dist.get_world_size())


class GraphedTrainStep:
    """One training step -- forward, loss, backward, gradient all-reduce, clipping, optimizer -- captured in a single
    CUDA graph and replayed.  A GKGNet-576 step is ~3000 kernel launches for 16 images; issued from Python they take
    longer (42 ms) than the kernels run (34 ms).  ``model(img, tgt)`` must return the scalar loss; the optimizer must be
    capturable (``torch.optim.AdamW(..., fused=True, capturable=True)``).  The gradient exchange is the same mean over
    ranks DDP computes, issued after the backward pass (138 MB over NVLink: < 1 ms, not worth overlapping)."""

    def __init__(self, model, optimizer, params, img, tgt, clip_norm: float | None = 5.0, warmup: int = 3,
                 amp_dtype=torch.bfloat16):
        if not img.is_cuda:
            raise ValueError("GraphedTrainStep needs CUDA tensors (there is no CPU path)")
        self.img, self.tgt = img.clone(), tgt.clone()
        # one all-reduce over the whole flat buffer: nothing overlaps it inside the graph, so a single large message
        # (best bus bandwidth) beats DDP-sized buckets (8 GPUs: 0.82 ms for six 25 MB buckets)
        self.grads = FlatGradients(params, bucket_bytes=1 << 40)
        params = self.grads.params

        views = [p.grad for p in params]                 # the flat buffer, parameter by parameter

        def body():
            # autograd writes fresh gradient tensors (p.grad is None: no accumulate kernel per parameter, no memset);
            # one multi-tensor copy then gathers them into the flat buffer, whose views become the gradients again
            for p in params:
                p.grad = None
            with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
                loss = model(self.img, self.tgt)
            loss.backward()
            fresh = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
            torch._foreach_copy_(views, fresh)
            for p, v in zip(params, views):
                p.grad = v
            del fresh
            self.grads.all_reduce_mean()
            if clip_norm is not None:
                torch.nn.utils.clip_grad_norm_(params, clip_norm)
            optimizer.step()
            return loss.detach()

        side = torch.cuda.Stream(device=img.device)
        side.wait_stream(torch.cuda.current_stream(img.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                body()
        torch.cuda.current_stream(img.device).wait_stream(side)
        torch.cuda.synchronize(img.device)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = body()
        self.launches_per_replay = _lib.launch_count() - n0      # libgkg_b200 kernels inside the graph

    def __call__(self, img=None, tgt=None):
        if img is not None:
            self.img.copy_(img, non_blocking=True)
        if tgt is not None:
            self.tgt.copy_(tgt, non_blocking=True)
        self.graph.replay()
        return self.loss

    # ---- input pipeline: the next batch crosses PCIe on a copy stream while the current step runs -------------------
    def prefetch(self, img_host, tgt_host):
        """Start the host -> device copy of the NEXT batch (pinned host tensors) into staging buffers on a copy stream."""
        dev = self.img.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = (torch.empty_like(self.img), torch.empty_like(self.tgt))
            self._ready, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        cs = self._copy_stream
        cs.wait_event(self._consumed)                      # the previous batch has left the staging buffers
        with torch.cuda.stream(cs):
            self._stage[0].copy_(img_host, non_blocking=True)
            self._stage[1].copy_(tgt_host, non_blocking=True)
            self._ready.record(cs)

    def step_prefetched(self):
        """Run one step on the batch handed to prefetch(): device -> device copy into the graph's static inputs, replay."""
        cur = torch.cuda.current_stream(self.img.device)
        cur.wait_event(self._ready)
        self.img.copy_(self._stage[0])
        self.tgt.copy_(self._stage[1])
        self._consumed.record(cur)
        self.graph.replay()
        return self.loss


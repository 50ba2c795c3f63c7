"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the
CPU tests).

* Forecasting shards by series and needs no collective (every series is independent: per-series statistics,
  causal attention within a series) — ``shard_range`` / ``shard_batch``.
* Fine-tuning the fusion module is data parallel; the only collective on the path is one all-reduce of the
  flattened fusion gradients per optimizer step (<= 30 MB, typically 1.97 MB), issued between ``backward`` and
  ``clip_grad_norm_`` so the clip sees the global-batch gradient (reference trainer.py:210-215 runs single-device).
  Full fine-tuning ("baseline" mode) all-reduces every adapter gradient the same way, in buckets of ``BUCKET_ELEMS``.
"""

from __future__ import annotations

import os
from collections.abc import Iterable

import torch
import torch.distributed as dist


def world_info() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not distributed."""
    return (
        int(os.environ.get("RANK", "0")),
        int(os.environ.get("WORLD_SIZE", "1")),
        int(os.environ.get("LOCAL_RANK", "0")),
    )


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    rank, world, local_rank = world_info()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of ``total`` series: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Slice every tensor / list of a collated batch along the series axis for this rank."""
    n = len(batch["context"])
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


BUCKET_ELEMS = 4 * 1024 * 1024  # floats per flattened bucket of SMALL tensors (16 MB)
DIRECT_ELEMS = 256 * 1024  # tensors of at least this many floats are all-reduced in place, without flattening


def allreduce_(tensors: Iterable[torch.Tensor], op: str = "mean", group=None) -> None:
    """In-place all-reduce ("sum" or "mean" over ranks) of a list of fp32 (gradient) tensors.

    Large contiguous tensors are reduced where they are; the small ones (norm scales, biases, the 80-float attention
    vectors: hundreds of them in a full fine-tune) travel in flattened buckets so that the step does not pay one
    collective launch per tensor.  No tensor is copied more than once each way and nothing the size of the whole model
    is ever allocated (the first version ``torch.cat``-ed 256 MB buckets: three extra passes over 1.99 GB)."""
    if op not in ("sum", "mean"):
        raise ValueError(f"op must be 'sum' or 'mean', got {op!r}")
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    works = []
    small: list[torch.Tensor] = []
    for t in tensors:
        if t.numel() >= DIRECT_ELEMS and t.is_contiguous() and t.dtype == torch.float32:
            works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
        else:
            small.append(t)
    flats: list[tuple[torch.Tensor, list[torch.Tensor]]] = []
    bucket: list[torch.Tensor] = []
    count = 0

    def flush() -> None:
        nonlocal bucket, count
        if bucket:
            flat = torch.cat([t.reshape(-1).to(torch.float32) for t in bucket])
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
            flats.append((flat, bucket))
            bucket, count = [], 0

    for t in small:
        if count and count + t.numel() > BUCKET_ELEMS:
            flush()
        bucket.append(t)
        count += t.numel()
    flush()
    for w in works:
        w.wait()
    for flat, members in flats:
        offset = 0
        for t in members:
            n = t.numel()
            t.copy_(flat[offset : offset + n].view_as(t))
            offset += n
    if op == "mean":
        torch._foreach_div_(tensors, float(world))


class OverlappedGradReducer:
    """Sum-all-reduces groups of gradient tensors as the backward pass hands them over (``MultimodalDecoder.
    grad_ready_hook``), so that the collective of layer i runs on NCCL's stream while the kernels of the layers below
    are still computing; ``finish()`` joins everything before the gradients are used.

    Large tensors are reduced in place, one asynchronous collective each, as soon as they arrive; the small ones (norm
    scales, biases, 80-float attention vectors) are collected and travel in one flattened bucket at ``finish()``.
    ``reduced`` tells the trainer that the gradients it finds in ``.grad`` are already global sums."""

    def __init__(self, group=None) -> None:
        self.group = group
        self._works: list = []
        self._small: list[torch.Tensor] = []
        self.reduced = False
        self.bytes = 0

    def __call__(self, tensors: Iterable[torch.Tensor]) -> None:
        for t in tensors:
            if t is None:
                continue
            self.bytes += t.numel() * t.element_size()
            if t.numel() >= DIRECT_ELEMS and t.is_contiguous() and t.dtype == torch.float32:
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self._small.append(t)

    def finish(self) -> None:
        small, self._small = self._small, []
        if small:
            flat = torch.cat([t.reshape(-1).to(torch.float32) for t in small])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            offset = 0
            for t in small:
                n = t.numel()
                t.copy_(flat[offset : offset + n].view_as(t))
                offset += n
        for w in self._works:
            w.wait()
        self._works = []
        self.reduced = True


def allreduce_mean_(tensors: Iterable[torch.Tensor], group=None) -> None:
    """In-place mean over ranks (see ``allreduce_``)."""
    allreduce_(tensors, "mean", group)


def allreduce_max(value: float, device: torch.device, group=None) -> float:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

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


BUCKET_ELEMS = 64 * 1024 * 1024  # floats per flattened all-reduce bucket (256 MB)


def allreduce_mean_(tensors: Iterable[torch.Tensor], group=None) -> None:
    """In-place mean over ranks of a list of (gradient) tensors through flattened all-reduces (one per bucket)."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    # flattened buckets of at most BUCKET_ELEMS floats: one collective for the fusion gradients (2 MB), a handful of
    # 256 MB ones for a full fine-tune (498 M parameters at 50 layers) without a second copy of all gradients at once
    bucket: list[torch.Tensor] = []
    count = 0

    def flush() -> None:
        nonlocal bucket, count
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        offset = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[offset : offset + n].view_as(t))
            offset += n
        bucket, count = [], 0

    for t in tensors:
        if count and count + t.numel() > BUCKET_ELEMS:
            flush()
        bucket.append(t)
        count += t.numel()
    flush()


def allreduce_max(value: float, device: torch.device, group=None) -> float:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

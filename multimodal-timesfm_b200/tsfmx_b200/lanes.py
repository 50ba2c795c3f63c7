"""Series lanes: overlap the tensor-bound and the HBM-bound kernels of the forecast path on one GPU.

Every series of a batch is independent on the whole path (per-series statistics, attention inside a series,
row-wise GEMMs), so a batch can be cut along the series axis into ``L`` *lanes* whose kernel chains never touch
each other's data.  Each lane gets its own CUDA stream and the host enqueues the lanes' launches round-robin, one
kernel at a time.  On the device the persistent tcgen05 GEMM of one lane (one CTA per SM, ~200 KB of shared memory,
issue slots almost idle) co-resides with the other lane's RMSNorm / attention / patchify CTAs (no shared memory to
speak of, HBM-bound), so a decoder layer costs about max(tensor time, HBM time) instead of their sum.  Results are
bit-identical to the single-stream order: no kernel's arithmetic depends on which other rows share its launch.

A stage that wants to take part exposes ``<stage>_steps(...)``, a generator that ``yield``s after every kernel
launch and ``return``s what ``<stage>(...)`` would; stages without one run as a single step.
"""

from __future__ import annotations

from collections.abc import Callable, Generator, Sequence
from typing import Any

import torch

_streams: dict[tuple[int, int], torch.cuda.Stream] = {}


def lane_streams(device: torch.device, count: int) -> list[torch.cuda.Stream]:
    index = device.index if device.index is not None else torch.cuda.current_device()
    out = []
    for i in range(count):
        key = (index, i)
        if key not in _streams:
            _streams[key] = torch.cuda.Stream(device=index)
        out.append(_streams[key])
    return out


def steps(obj: Any, name: str, *args: Any) -> Generator[None, None, Any]:
    """``yield from steps(adapter, "forward", x, m)``: the stage's own step generator if it has one."""
    gen_fn = getattr(obj, name + "_steps", None)
    if gen_fn is not None:
        return (yield from gen_fn(*args))
    result = getattr(obj, name)(*args)
    yield
    return result


def drain(gen: Generator[None, None, Any]) -> Any:
    """Run a step generator to completion on the current stream."""
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        return stop.value


def run_lanes(make_gen: Callable[[int], Generator[None, None, Any]], count: int, device: torch.device) -> list[Any]:
    """Run ``count`` step generators, lane ``i`` on its own stream, launches interleaved round-robin.

    The lane streams first wait for everything already enqueued on the caller's stream (the inputs), and the
    caller's stream waits for all lanes before this returns, so the call is stream-ordered like a plain kernel.
    Tensors a lane returns were allocated on the lane's stream; they are handed to the caller's stream with
    ``record_stream`` so that the caching allocator does not recycle them early.
    """
    if device.index is not None and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):  # set_stream() below selects a stream, not a device
            return run_lanes(make_gen, count, device)
    main = torch.cuda.current_stream(device)
    streams = lane_streams(device, count)
    for s in streams:
        s.wait_stream(main)
    gens: list[Generator[None, None, Any] | None] = []
    results: list[Any] = [None] * count
    try:
        for i in range(count):
            torch.cuda.set_stream(streams[i])
            gens.append(make_gen(i))
        live = list(range(count))
        while live:
            for i in list(live):
                torch.cuda.set_stream(streams[i])
                try:
                    next(gens[i])
                except StopIteration as stop:
                    results[i] = stop.value
                    live.remove(i)
    finally:
        torch.cuda.set_stream(main)
    for s in streams:
        main.wait_stream(s)
    for r in results:
        for t in _tensors(r):
            t.record_stream(main)
    return results


def _tensors(obj: Any):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors(v)


def split_points(batch: int, count: int, multiple: int = 1) -> Sequence[tuple[int, int]]:
    """Contiguous [lo, hi) ranges of the series axis, sizes rounded to ``multiple`` where possible."""
    count = max(1, min(count, batch))
    base = -(-batch // count)
    if multiple > 1:
        base = -(-base // multiple) * multiple
    out, lo = [], 0
    while lo < batch:
        hi = min(batch, lo + base)
        out.append((lo, hi))
        lo = hi
    return out

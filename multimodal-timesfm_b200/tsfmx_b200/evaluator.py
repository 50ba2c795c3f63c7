"""Forecast entry loop (reference tsfmx/evaluator.py:12-71): no-grad pass over a loader -> sample-weighted MSE / MAE.

Same class, constructor and ``evaluate`` contract as the reference.  Differences that do not change results:

* host batches are staged one batch ahead on a copy stream (``non_blocking`` copies from pinned memory), so the
  host-to-device transfer of batch i + 1 - the text embeddings are ten times the series bytes - overlaps the
  kernels of batch i;
* the per-batch ``.item()`` host syncs of the reference (evaluator.py:61-62) become asynchronous 16-byte read-backs
  into pinned memory that are summed once at the end;
* the staging buffers live as long as the evaluator, and the forecast of a staged batch is replayed from a CUDA graph
  (one per staging slot, ``MultimodalDecoder._graphed_forecast``), so the host's per-launch Python cost is off the
  step (``graphs=False`` keeps the eager launches).
"""

from __future__ import annotations

from collections.abc import Iterable, Iterator
from typing import TypedDict

import torch

from .decoder import MultimodalDecoder


class EvaluationMetrics(TypedDict):
    mse: float
    mae: float


class MultimodalEvaluator:
    """Computes evaluation metrics for a multimodal decoder (text embeddings are fused when the batch has them)."""

    _KEYS = ("context", "horizon", "text_embeddings")
    MAX_CAPTURES_PER_PASS = 6  # two staging slots x (full batch, ragged tail) and some slack

    def __init__(self, model: MultimodalDecoder, device: torch.device, *, graphs: bool | None = None) -> None:
        self.model = model
        self.device = torch.device(device)
        self.graphs = self.device.type == "cuda" if graphs is None else bool(graphs)
        # device staging slots (and their padding masks), reused across evaluate() calls: stable addresses are what
        # lets the forecast graphs be replayed
        self._slots: list[dict[str, torch.Tensor]] = [{}, {}]
        self._copy_stream: torch.cuda.Stream | None = None
        self._d2h_stream: torch.cuda.Stream | None = None

    def _staged(self, dataloader: Iterable[dict]) -> Iterator[dict]:
        """Yield device-resident batches; with a CUDA device the next batch is already in flight on a copy stream."""
        if self.device.type != "cuda":
            for batch in dataloader:
                out = {k: batch[k].to(self.device) for k in self._KEYS if k in batch}
                out["input_padding"] = torch.zeros_like(out["context"], dtype=torch.bool)  # reference evaluator.py:52
                yield out
            return
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy = self._copy_stream
        copy.wait_stream(main)  # kernels of an earlier pass may still be reading the slots
        # two sets of device staging buffers, reused for the evaluator's lifetime (no allocator traffic per batch): slot
        # s is refilled on the copy stream only after the kernels that consumed its previous contents have finished
        slots = self._slots
        consumed: list[torch.cuda.Event | None] = [None, None]

        def stage(batch: dict, slot: int):
            with torch.cuda.stream(copy):
                if consumed[slot] is not None:
                    copy.wait_event(consumed[slot])
                out = {}
                for k in self._KEYS:
                    if k not in batch:
                        continue
                    src = batch[k]
                    # the slot keeps its largest batch: a ragged last batch is a leading view of the same storage, so
                    # the addresses (and the graphs bound to them) survive from pass to pass
                    full = slots[slot].get(k)
                    if (full is None or full.shape[1:] != src.shape[1:] or full.dtype != src.dtype
                            or full.shape[0] < src.shape[0]):
                        full = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                        full.record_stream(main)
                        slots[slot][k] = full
                    buf = full[: src.shape[0]]
                    buf.copy_(src, non_blocking=True)
                    out[k] = buf
                pad = slots[slot].get("input_padding")
                want = out["context"].shape
                if pad is None or pad.shape[1:] != want[1:] or pad.shape[0] < want[0]:
                    pad = torch.zeros(want, dtype=torch.bool, device=self.device)
                    pad.record_stream(main)
                    slots[slot]["input_padding"] = pad
                out["input_padding"] = pad[: want[0]]
                done = torch.cuda.Event()
                done.record(copy)
            return out, done

        it = iter(dataloader)
        try:
            pending = stage(next(it), 0)
        except StopIteration:
            return
        index = 0
        while pending is not None:
            cur, done = pending
            try:
                pending = stage(next(it), (index + 1) % 2)
            except StopIteration:
                pending = None
            main.wait_event(done)
            yield cur
            # the consumer has enqueued everything that reads this slot
            consumed[index % 2] = torch.cuda.Event()
            consumed[index % 2].record(main)
            index += 1

    def predict(self, dataloader: Iterable[dict], horizon: int | None = None, *, full: bool = True,
                copy: bool = True) -> Iterator[torch.Tensor]:
        """Forecasts of every batch of ``dataloader`` as HOST tensors: ``(B, horizon, Q)`` quantile forecasts
        (``full``) or the ``(B, horizon)`` point forecast.  Not in the reference (its only consumer of forecasts is
        ``evaluate``); this is the same staged loop for callers that want the numbers themselves.

        The inputs are staged one batch ahead like in ``evaluate``; every forecast is copied device -> host on its own
        stream into one of three page-locked buffers while the next batch computes.  ``copy=False`` yields views of
        those buffers: each stays valid until two more batches have been requested.  ``horizon`` defaults to the
        length of the batch's ``"horizon"`` target."""
        self.model.eval()
        cuda = self.device.type == "cuda"
        graphs_before = getattr(self.model, "graphs", False)
        if cuda and hasattr(self.model, "graphs"):
            self.model.graphs = self.graphs
        try:
            yield from self._predict(dataloader, horizon, full, copy, cuda)
        finally:
            if hasattr(self.model, "graphs"):
                self.model.graphs = graphs_before

    def _predict(self, dataloader, horizon, full, copy, cuda) -> Iterator[torch.Tensor]:
        fn = self.model.forward_full if full else self.model
        ring: list[torch.Tensor | None] = [None, None, None]
        copied: list[torch.cuda.Event | None] = [None, None, None]
        pending: list[tuple[int, tuple[int, ...]]] = []  # (ring slot, shape) of forecasts whose read-back is in flight
        main = torch.cuda.current_stream(self.device) if cuda else None
        if cuda and self._d2h_stream is None:
            self._d2h_stream = torch.cuda.Stream(device=self.device)
        d2h = self._d2h_stream

        def land(slot: int, shape: tuple[int, ...]) -> torch.Tensor:
            copied[slot].synchronize()
            view = ring[slot][: shape[0]]
            return view.clone() if copy else view

        with torch.no_grad():
            for index, batch in enumerate(self._staged(dataloader)):
                context = batch["context"]
                h = int(horizon if horizon is not None else batch["horizon"].shape[-1])
                if not cuda:
                    yield fn(h, context, batch["input_padding"], batch.get("text_embeddings"))
                    continue
                slot = index % 3
                if index >= 2:
                    # the staging slot this batch sits in (index % 2) fed the forecast of batch index - 2, whose
                    # output buffer a graph replay is about to overwrite: its read-back must have finished
                    main.wait_event(copied[(index - 2) % 3])
                out = fn(h, context, batch["input_padding"], batch.get("text_embeddings"))
                if ring[slot] is None or ring[slot].shape[1:] != out.shape[1:] or ring[slot].shape[0] < out.shape[0]:
                    ring[slot] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                computed = torch.cuda.Event()
                computed.record(main)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(computed)
                    ring[slot][: out.shape[0]].copy_(out, non_blocking=True)
                    out.record_stream(d2h)
                    copied[slot] = torch.cuda.Event()
                    copied[slot].record(d2h)
                pending.append((slot, tuple(out.shape)))
                if len(pending) == 2:  # hand out batch i - 1 while batch i computes
                    yield land(*pending.pop(0))
        for slot, shape in pending:
            yield land(slot, shape)

    def evaluate(self, dataloader: Iterable[dict]) -> EvaluationMetrics:
        """Raises RuntimeError if the loader yields no samples (reference evaluator.py:65-66)."""
        self.model.eval()
        cuda = self.device.type == "cuda"
        graphs_before = getattr(self.model, "graphs", False)
        if cuda and hasattr(self.model, "graphs"):
            self.model.graphs = self.graphs
        try:
            return self._evaluate(dataloader, cuda)
        finally:
            if hasattr(self.model, "graphs"):
                self.model.graphs = graphs_before

    def _evaluate(self, dataloader: Iterable[dict], cuda: bool) -> EvaluationMetrics:
        sums: list[torch.Tensor] = []   # per batch [mean squared error, mean absolute error] * n
        ring = torch.empty(256, 2, dtype=torch.float64, pin_memory=True) if cuda else None  # one pinned allocation
        used = 0
        num_samples = 0
        captures_at_start = getattr(self.model, "graph_captures", 0)
        with torch.no_grad():
            for batch in self._staged(dataloader):
                context, horizon = batch["context"], batch["horizon"]
                # a loader whose batch shapes keep changing would capture a graph (two forwards) per batch: after
                # MAX_CAPTURES_PER_PASS of them the rest of the pass runs the eager launches
                if getattr(self.model, "graph_captures", 0) - captures_at_start >= self.MAX_CAPTURES_PER_PASS:
                    self.model.graphs = False
                # all-False padding mask (reference evaluator.py:52), one per staging slot
                point = self.model(horizon.shape[-1], context, batch["input_padding"], batch.get("text_embeddings"))
                err = point - horizon
                n = context.size(0)
                stat = torch.stack([err.square().mean(), err.abs().mean()]).double() * n
                if cuda:
                    if used == ring.shape[0]:  # ring full: fold what has landed into one row
                        torch.cuda.current_stream(self.device).synchronize()
                        sums.append(ring.sum(0))
                        used = 0
                    ring[used].copy_(stat, non_blocking=True)  # 16 bytes per batch, no sync
                    used += 1
                else:
                    sums.append(stat)
                num_samples += n
        if num_samples == 0:
            raise RuntimeError("Evaluation dataset is empty.")
        if cuda:
            torch.cuda.current_stream(self.device).synchronize()
            sums.append(ring[:used].sum(0))
        total = torch.stack(sums).sum(0)
        mse, mae = (total / num_samples).tolist()
        return EvaluationMetrics(mse=mse, mae=mae)

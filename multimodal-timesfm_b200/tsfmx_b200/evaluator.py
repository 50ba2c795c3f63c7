"""Forecast entry loop (reference tsfmx/evaluator.py:12-71): no-grad pass over a loader -> sample-weighted MSE / MAE.

Same class, constructor and ``evaluate`` contract as the reference.  Differences that do not change results:

* host batches are staged one batch ahead on a copy stream (``non_blocking`` copies from pinned memory), so the
  host-to-device transfer of batch i + 1 - the text embeddings are ten times the series bytes - overlaps the
  kernels of batch i;
* the per-batch ``.item()`` host syncs of the reference (evaluator.py:61-62) become asynchronous 16-byte read-backs
  into pinned memory that are summed once at the end.
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

    def __init__(self, model: MultimodalDecoder, device: torch.device) -> None:
        self.model = model
        self.device = torch.device(device)

    def _staged(self, dataloader: Iterable[dict]) -> Iterator[dict]:
        """Yield device-resident batches; with a CUDA device the next batch is already in flight on a copy stream."""
        if self.device.type != "cuda":
            for batch in dataloader:
                yield {k: batch[k].to(self.device) for k in self._KEYS if k in batch}
            return
        main = torch.cuda.current_stream(self.device)
        copy = torch.cuda.Stream(device=self.device)
        # two sets of device staging buffers, reused for the whole pass (no allocator traffic per batch): slot s is
        # refilled on the copy stream only after the kernels that consumed its previous contents have finished
        slots: list[dict[str, torch.Tensor]] = [{}, {}]
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
                    buf = slots[slot].get(k)
                    if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
                        buf = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                        buf.record_stream(main)
                        slots[slot][k] = buf
                    buf.copy_(src, non_blocking=True)
                    out[k] = buf
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

    def evaluate(self, dataloader: Iterable[dict]) -> EvaluationMetrics:
        """Raises RuntimeError if the loader yields no samples (reference evaluator.py:65-66)."""
        self.model.eval()
        cuda = self.device.type == "cuda"
        sums: list[torch.Tensor] = []   # per batch [mean squared error, mean absolute error] * n
        ring = torch.empty(256, 2, dtype=torch.float64, pin_memory=True) if cuda else None  # one pinned allocation
        used = 0
        num_samples = 0
        with torch.no_grad():
            for batch in self._staged(dataloader):
                context, horizon = batch["context"], batch["horizon"]
                input_padding = torch.zeros_like(context, dtype=torch.bool)  # reference evaluator.py:52
                point = self.model(horizon.shape[-1], context, input_padding, batch.get("text_embeddings"))
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

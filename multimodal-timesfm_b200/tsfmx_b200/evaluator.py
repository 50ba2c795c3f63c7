"""Forecast entry loop (reference tsfmx/evaluator.py:12-71): no-grad pass over a loader -> sample-weighted MSE / MAE.

Same class, constructor and ``evaluate`` contract as the reference.  Differences that do not change results:
the per-batch ``.item()`` host syncs of the reference (evaluator.py:61-62) are replaced by on-device accumulation
with a single read-back at the end, and host batches are copied with ``non_blocking=True`` (pinned loaders overlap the
copy with the previous batch's kernels).
"""

from __future__ import annotations

from collections.abc import Iterable
from typing import TypedDict

import torch

from .decoder import MultimodalDecoder


class EvaluationMetrics(TypedDict):
    mse: float
    mae: float


class MultimodalEvaluator:
    """Computes evaluation metrics for a multimodal decoder (text embeddings are fused when the batch has them)."""

    def __init__(self, model: MultimodalDecoder, device: torch.device) -> None:
        self.model = model
        self.device = device

    def evaluate(self, dataloader: Iterable[dict]) -> EvaluationMetrics:
        """Raises RuntimeError if the loader yields no samples (reference evaluator.py:65-66)."""
        self.model.eval()
        total = torch.zeros(2, dtype=torch.float64, device=self.device)
        num_samples = 0
        with torch.no_grad():
            for batch in dataloader:
                context = batch["context"].to(self.device, non_blocking=True)
                horizon = batch["horizon"].to(self.device, non_blocking=True)
                horizon_len = horizon.shape[-1]
                input_padding = torch.zeros_like(context, dtype=torch.bool)
                text = batch["text_embeddings"].to(self.device, non_blocking=True) if "text_embeddings" in batch else None
                point = self.model(horizon_len, context, input_padding, text)
                err = point - horizon
                n = context.size(0)
                total += torch.stack([err.square().mean(), err.abs().mean()]).double() * n
                num_samples += n
        if num_samples == 0:
            raise RuntimeError("Evaluation dataset is empty.")
        mse, mae = (total / num_samples).tolist()
        return EvaluationMetrics(mse=mse, mae=mae)

"""Collate functions (reference tsfmx/data/collate.py:9-29).  Same keys as the reference ``Batch`` TypedDict
(``context``, ``horizon``, optional ``text_embeddings``, ``metadata``; reference types.py:33-39)."""

from __future__ import annotations

from typing import Any

import numpy as np
import torch


def _build_batch(batch: list[dict[str, Any]]) -> dict[str, Any]:
    return {
        "context": torch.from_numpy(np.stack([s["context"] for s in batch])),
        "horizon": torch.from_numpy(np.stack([s["horizon"] for s in batch])),
        "metadata": [s["metadata"] for s in batch],
    }


def multimodal_collate_fn(batch: list[dict[str, Any]]) -> dict[str, Any]:
    """Batches with pre-computed per-patch text embeddings (reference collate.py:20-24)."""
    out = _build_batch(batch)
    out["text_embeddings"] = torch.from_numpy(np.stack([s["text_embeddings"] for s in batch]))
    return out


def baseline_collate_fn(batch: list[dict[str, Any]]) -> dict[str, Any]:
    """Batches without text (reference collate.py:27-29)."""
    return _build_batch(batch)

"""Packed, pinned sample store: the data path into the GPU.

The reference keeps a ``list[PreprocessedSample]`` (one dict of small numpy arrays per sample, pickled by
``PreprocessPipeline``, reference tsfmx/data/preprocess.py:60-72) and builds every batch with ``np.stack`` in
``collate_fn`` (reference tsfmx/data/collate.py:9-29).  At 50 k series/s per GPU that is 2.8 GB/s of host traffic in
24.6 KB pieces, so here the list is packed ONCE into three contiguous page-locked tensors - context ``[S, C]``, horizon
``[S, h]``, text embeddings ``[S, N, E]`` - and a batch is a set of zero-copy slices of them: the evaluator's copy
stream can DMA straight out of pinned memory, and shuffled epochs use one index_select per batch into a page-locked
buffer taken from torch's caching host allocator (which recycles a block only after every asynchronous copy that
read it has completed, so a consumer may run any number of batches ahead of the device).
Batches carry the same keys as the reference ``Batch`` TypedDict (types.py:33-39).
"""

from __future__ import annotations

import pickle
from collections.abc import Iterator, Sequence
from pathlib import Path
from typing import Any

import numpy as np
import torch


def _pinned(shape: tuple[int, ...], dtype: torch.dtype, pin: bool) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, pin_memory=pin)


class PackedSamples:
    """Contiguous (optionally page-locked) storage of preprocessed samples + batch iteration."""

    def __init__(self, context: torch.Tensor, horizon: torch.Tensor, text_embeddings: torch.Tensor | None,
                 metadata: list[dict[str, Any]]) -> None:
        if horizon.shape[0] != context.shape[0] or (text_embeddings is not None and text_embeddings.shape[0] != context.shape[0]):
            raise ValueError("context, horizon and text_embeddings must hold the same number of samples")
        self.context, self.horizon, self.text_embeddings, self.metadata = context, horizon, text_embeddings, metadata

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_samples(cls, samples: Sequence[dict[str, Any]], pin: bool | None = None) -> "PackedSamples":
        """Pack a reference-style sample list.  Every sample must have the same context / horizon / text shape (the
        reference's datasets produce fixed windows).  Raises ValueError on an empty or ragged list."""
        if len(samples) == 0:
            raise ValueError("cannot pack an empty sample list")
        pin = torch.cuda.is_available() if pin is None else pin
        first = samples[0]
        c_shape, h_shape = np.shape(first["context"]), np.shape(first["horizon"])
        has_text = "text_embeddings" in first
        t_shape = np.shape(first["text_embeddings"]) if has_text else None
        n = len(samples)
        context = _pinned((n, *c_shape), torch.float32, pin)
        horizon = _pinned((n, *h_shape), torch.float32, pin)
        text = _pinned((n, *t_shape), torch.float32, pin) if has_text else None
        cn, hn = context.numpy(), horizon.numpy()
        tn = text.numpy() if text is not None else None
        for i, s in enumerate(samples):
            if np.shape(s["context"]) != c_shape or np.shape(s["horizon"]) != h_shape or (
                has_text and np.shape(s.get("text_embeddings")) != t_shape
            ):
                raise ValueError(f"sample {i} does not have the shape of sample 0; ragged windows cannot be packed")
            cn[i], hn[i] = s["context"], s["horizon"]
            if tn is not None:
                tn[i] = s["text_embeddings"]
        return cls(context, horizon, text, [s.get("metadata", {}) for s in samples])

    @classmethod
    def from_pickle(cls, path: str | Path, pin: bool | None = None) -> "PackedSamples":
        """Load a cache file written by the reference's ``PreprocessPipeline`` (a pickled ``list[PreprocessedSample]``)."""
        with open(path, "rb") as f:
            samples = pickle.load(f)
        return cls.from_samples(samples, pin)

    # ------------------------------------------------------------------ access
    def __len__(self) -> int:
        return int(self.context.shape[0])

    def __getitem__(self, i: int) -> dict[str, Any]:
        """One sample in the reference's ``PreprocessedSample`` form (numpy views, no copy)."""
        out = {"context": self.context[i].numpy(), "horizon": self.horizon[i].numpy(), "metadata": self.metadata[i]}
        if self.text_embeddings is not None:
            out["text_embeddings"] = self.text_embeddings[i].numpy()
        return out

    def _batch(self, lo: int, hi: int) -> dict[str, Any]:
        out = {"context": self.context[lo:hi], "horizon": self.horizon[lo:hi], "metadata": self.metadata[lo:hi]}
        if self.text_embeddings is not None:
            out["text_embeddings"] = self.text_embeddings[lo:hi]
        return out

    def _gather(self, idx: torch.Tensor) -> dict[str, Any]:
        out: dict[str, Any] = {"metadata": [self.metadata[i] for i in idx.tolist()]}
        for key in ("context", "horizon", "text_embeddings"):
            src = getattr(self, key)
            if src is None:
                continue
            # a FRESH page-locked tensor per batch, never a reused scratch: a consumer that stages batches with
            # ``copy_(non_blocking=True)`` (MultimodalEvaluator, MultimodalTrainer) only ENQUEUES the copy before it
            # asks for the next batch, so a shared buffer would be overwritten by this gather while the DMA of the
            # previous batch is still pending.  torch's caching host allocator records the copying stream on the
            # block and hands it out again only once that copy has finished, so steady state allocates nothing.
            dst = _pinned((idx.numel(), *src.shape[1:]), src.dtype, src.is_pinned())
            torch.index_select(src, 0, idx, out=dst)
            out[key] = dst
        return out

    def batches(self, batch_size: int, shuffle: bool = False, generator: torch.Generator | None = None,
                drop_last: bool = False) -> Iterator[dict[str, Any]]:
        """Batches with the reference's ``Batch`` keys.  In order: zero-copy slices of the pinned store.  Shuffled: one
        gather per batch into its own page-locked tensor (recycled by torch's caching host allocator once the copies
        that read it are done), so the consumer may hold or be copying any number of earlier batches."""
        if batch_size < 1:
            raise ValueError(f"batch_size must be >= 1, got {batch_size}")
        n = len(self)
        if not shuffle:
            for lo in range(0, n, batch_size):
                hi = min(n, lo + batch_size)
                if drop_last and hi - lo < batch_size:
                    return
                yield self._batch(lo, hi)
            return
        perm = torch.randperm(n, generator=generator)
        for lo in range(0, n, batch_size):
            idx = perm[lo : lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield self._gather(idx)

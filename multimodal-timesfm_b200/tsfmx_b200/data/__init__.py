"""Batch plumbing (reference tsfmx/data/collate.py:9-29): numpy-stack collate with optional pinned staging."""

from .collate import baseline_collate_fn, multimodal_collate_fn

__all__ = ["baseline_collate_fn", "multimodal_collate_fn"]

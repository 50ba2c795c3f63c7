"""Host-side data helpers: the reference's collate functions and the packed, pinned sample store."""

from .collate import baseline_collate_fn, multimodal_collate_fn
from .packed import PackedSamples

__all__ = ["baseline_collate_fn", "multimodal_collate_fn", "PackedSamples"]

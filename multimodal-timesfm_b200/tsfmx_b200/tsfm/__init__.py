"""TSFM adapters (reference tsfmx/tsfm/)."""

from .base import PreprocessResult, TsfmAdapter

__all__ = ["PreprocessResult", "TsfmAdapter"]

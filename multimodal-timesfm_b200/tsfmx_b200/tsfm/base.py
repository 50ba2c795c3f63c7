"""Adapter plugin interface — the drop-in boundary of the hot path.

Same contract as the reference ``tsfmx.tsfm.base`` (reference tsfmx/tsfm/base.py:10-75): a
``PreprocessResult`` record and the ``TsfmAdapter`` ABC whose three stages bracket the fusion
injection point::

    preprocess -> [fusion] -> forward -> postprocess
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import torch
from torch import nn


@dataclass
class PreprocessResult:
    """What ``TsfmAdapter.preprocess`` hands to the fusion stage (reference base.py:10-22).

    input_embeddings: (batch, num_patches, model_dims) patch-token embeddings.
    masks: adapter-specific boolean mask, True = padded.
    normalization_stats: tensors needed to undo the input normalisation in ``postprocess``.
    """

    input_embeddings: torch.Tensor
    masks: torch.Tensor
    normalization_stats: dict[str, torch.Tensor]


class TsfmAdapter(nn.Module, ABC):
    """A time-series foundation model behind the three-stage adapter API (reference base.py:25-75)."""

    @property
    @abstractmethod
    def model_dims(self) -> int:
        """Width of the transformer's residual stream."""

    @property
    @abstractmethod
    def patch_len(self) -> int:
        """Raw time steps per input patch."""

    @property
    @abstractmethod
    def point_forecast_index(self) -> int:
        """Channel of the ``postprocess`` output that is the point forecast."""

    @abstractmethod
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult: ...

    @abstractmethod
    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor: ...

    @abstractmethod
    def postprocess(
        self,
        horizon: int,
        output_embeddings: torch.Tensor,
        normalization_stats: dict[str, torch.Tensor],
    ) -> torch.Tensor: ...

    @abstractmethod
    def freeze_parameters(self) -> None: ...

    @abstractmethod
    def unfreeze_parameters(self) -> None: ...

"""Adapter plugin interface — the drop-in boundary of the hot path.

Same contract as the reference ``tsfmx.tsfm.base`` (reference tsfmx/tsfm/base.py:10-75): a
``PreprocessResult`` record and the ``TsfmAdapter`` ABC whose three stages bracket the fusion
injection point::

    preprocess -> [fusion] -> forward -> postprocess

On top of the reference's abstract members the base class names the OPTIONAL hooks the B200 host code looks
for (series lanes, CUDA-graph replay, the hand-written backward): an adapter that does not override them runs
the plain three-call path, exactly like an adapter written against the reference.
"""

from __future__ import annotations

import abc
import dataclasses

import torch
from torch import nn

Stats = dict[str, torch.Tensor]


@dataclasses.dataclass
class PreprocessResult:
    """What ``TsfmAdapter.preprocess`` hands to the fusion stage (reference base.py:10-22)."""

    input_embeddings: torch.Tensor  # (batch, num_patches, model_dims) patch-token embeddings
    masks: torch.Tensor  # adapter-specific boolean mask, True = padded
    normalization_stats: Stats  # whatever ``postprocess`` needs to undo the input normalisation


class TsfmAdapter(nn.Module, abc.ABC):
    """A time-series foundation model behind the three-stage adapter API (reference base.py:25-75)."""

    # ------------------------------------------------------------------ the three stages
    @abc.abstractmethod
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult:
        """(batch, context) series + boolean padding mask -> patch-token embeddings, masks, statistics."""

    @abc.abstractmethod
    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """The transformer stack on (possibly fused) embeddings -> output embeddings."""

    @abc.abstractmethod
    def postprocess(self, horizon: int, output_embeddings: torch.Tensor, normalization_stats: Stats) -> torch.Tensor:
        """Output head + inverse normalisation -> (batch, horizon, num_outputs) forecasts."""

    # ------------------------------------------------------------------ what the decoder / trainer ask about the model
    @property
    @abc.abstractmethod
    def model_dims(self) -> int:
        """Width of the transformer's residual stream (= what the fusion projection must produce)."""

    @property
    @abc.abstractmethod
    def patch_len(self) -> int:
        """Raw time steps per input patch (the datasets cut their per-patch text embeddings with it)."""

    @property
    @abc.abstractmethod
    def point_forecast_index(self) -> int:
        """Channel of the ``postprocess`` output that ``MultimodalDecoder.forward`` returns."""

    @abc.abstractmethod
    def freeze_parameters(self) -> None:
        """``requires_grad = False`` on every parameter ("multimodal" training mode, trainer.py:76-77)."""

    @abc.abstractmethod
    def unfreeze_parameters(self) -> None:
        """``requires_grad = True`` on every parameter ("baseline" training mode, trainer.py:78-79)."""

    # ------------------------------------------------------------------ optional hooks of the B200 host code
    #: number of forecast channels (lets the decoder return an empty forecast for an empty batch)
    num_outputs: int | None = None

    @property
    def graph_safe(self) -> bool:
        """True if the three stages make no host synchronisation, so ``MultimodalDecoder`` may capture them into a
        CUDA graph.  Adapters may additionally provide ``preprocess_steps`` / ``forward_steps`` /
        ``postprocess_steps`` generators (one yield per kernel launch, ``tsfmx_b200.lanes``) and, for training,
        ``forward_saving`` / ``forward_backward`` / ``postprocess_saving`` / ``postprocess_backward`` (fusion fine-tuning)
        plus ``preprocess_saving`` / ``preprocess_backward`` (full fine-tuning)."""
        return False

"""Chronos-T5 tokeniser (BASELINE.json north_star: "Chronos mean-scaling plus bin quantisation", bit-exact ids).

Not part of the reference (it only wraps Chronos-2); provided as the extra plugin stage SURVEY.md section 9 asks for,
with the upstream ``chronos.MeanScaleUniformBins`` method names.  Both transforms are single fused CUDA kernels
(``tsfmx_chronos_t5_tokenize`` / ``tsfmx_chronos_t5_dequantize``); the T5 encoder-decoder backbone itself is not
rebuilt here (listed under "next" in DESIGN.md).
"""

from __future__ import annotations

import torch
from torch import nn

from .. import ops


class MeanScaleUniformBins(nn.Module):
    """chronos-t5-* tokenizer: n_tokens 4096, 2 special tokens (PAD 0, EOS 1), bin centres linspace(-15, 15, 4093)."""

    def __init__(self, n_tokens: int = 4096, n_special_tokens: int = 2, low_limit: float = -15.0,
                 high_limit: float = 15.0, pad_token_id: int = 0, eos_token_id: int = 1) -> None:
        super().__init__()
        self.n_tokens, self.n_special_tokens = n_tokens, n_special_tokens
        self.pad_token_id, self.eos_token_id = pad_token_id, eos_token_id
        centers = torch.linspace(low_limit, high_limit, n_tokens - n_special_tokens - 1)
        boundaries = torch.concat((torch.tensor([-1e20]), (centers[1:] + centers[:-1]) / 2, torch.tensor([1e20])))
        self.register_buffer("centers", centers)
        self.register_buffer("boundaries", boundaries)

    def context_input_transform(self, context: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """context (B, C) fp32, NaN = missing -> token ids (B, C+1) int64 with EOS, attention mask (B, C+1), scale (B,)."""
        return ops.chronos_t5_tokenize(context, self.boundaries, self.n_special_tokens, self.n_tokens,
                                       self.pad_token_id, self.eos_token_id)

    def output_transform(self, samples: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
        """token ids (B, L) -> values = centers[clamp(id - n_special - 1)] * scale."""
        return ops.chronos_t5_dequantize(samples, self.centers, scale, self.n_special_tokens)

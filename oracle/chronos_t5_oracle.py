"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/timesfm_oracle.py header for the import rules).

Chronos-T5 ``MeanScaleUniformBins`` tokeniser (BASELINE.json north_star item; NOT part of the reference, which
only wraps Chronos-2).  Restates the published algorithm of ``chronos-forecasting``
(``chronos.chronos.MeanScaleUniformBins._input_transform`` / ``_append_eos_token`` / ``output_transform``,
SURVEY.md appendix A.3) with the `chronos-t5-*` configuration: n_tokens 4096, n_special_tokens 2 (PAD 0, EOS 1),
low/high limit -15/+15, use_eos_token.

PARITY PINNING: upstream source is not available in this container -> "parity unpinned" against upstream; the
restatement is pinned against torch's own ``bucketize`` / ``linspace`` (tests/test_oracle_cpu.py).

One deliberate, documented choice: upstream computes ``scale = nansum(|x|) / nansum(mask)`` with an fp32
``torch.nansum`` whose rounding depends on the reduction order (CPU vector width, CUDA block shape), i.e. the
ids it produces are not reproducible across machines at bin edges.  The oracle (and the CUDA kernel) define
``sum(|x|)`` as the exactly-accumulated (fp64) sum rounded once to fp32, which is order independent and is
within a few ulp of any fp32 summation order; everything after it (fp32 division, bucketize right=True, +2,
clamp, PAD for NaN, EOS) is bit-for-bit the upstream arithmetic.
"""

from __future__ import annotations

import torch

N_TOKENS, N_SPECIAL, PAD_ID, EOS_ID = 4096, 2, 0, 1
LOW, HIGH = -15.0, 15.0


def tables(n_tokens: int = N_TOKENS, n_special: int = N_SPECIAL, low: float = LOW, high: float = HIGH):
    centers = torch.linspace(low, high, n_tokens - n_special - 1)
    boundaries = torch.concat(
        (torch.tensor([-1e20]), (centers[1:] + centers[:-1]) / 2, torch.tensor([1e20]))
    )
    return centers, boundaries


def tokenize(context: torch.Tensor, boundaries: torch.Tensor, n_tokens: int = N_TOKENS, n_special: int = N_SPECIAL):
    """context (B, C) fp32 with NaN = missing -> ids (B, C+1) int64, attention_mask (B, C+1) bool, scale (B,) fp32."""
    context = context.to(torch.float32)
    attention_mask = ~torch.isnan(context)
    abs_sum = torch.nansum(torch.abs(context).double() * attention_mask, dim=-1).to(torch.float32)
    scale = abs_sum / torch.nansum(attention_mask.to(torch.float32), dim=-1)
    scale[~(scale > 0)] = 1.0
    scaled = context / scale.unsqueeze(-1)
    ids = torch.bucketize(scaled, boundaries, right=True) + n_special
    ids.clamp_(0, n_tokens - 1)
    ids[~attention_mask] = PAD_ID
    b = context.shape[0]
    ids = torch.concat((ids, torch.full((b, 1), EOS_ID, dtype=ids.dtype)), dim=1)
    attention_mask = torch.concat((attention_mask, torch.ones(b, 1, dtype=torch.bool)), dim=1)
    return ids, attention_mask, scale


def dequantize(ids: torch.Tensor, centers: torch.Tensor, scale: torch.Tensor, n_special: int = N_SPECIAL):
    idx = torch.clamp(ids - n_special - 1, 0, len(centers) - 1)
    return centers[idx] * scale.unsqueeze(-1)

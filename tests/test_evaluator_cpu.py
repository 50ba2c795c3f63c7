"""Host logic of the forecast entry loop (tsfmx_b200.evaluator) on the CPU with a stub model, pinned to the reference's
own MultimodalEvaluator (imported from /root/reference/src when present): same metrics, same RuntimeError."""

import sys
from pathlib import Path

import pytest
import torch
from torch import nn

from tsfmx_b200.evaluator import MultimodalEvaluator

REF = Path("/root/reference/src")


class _Stub(nn.Module):
    """forecast = last context value + 0.1 * mean text embedding (or 0), repeated over the horizon."""

    def forward(self, horizon, inputs, masks, text_embeddings=None):
        assert masks.dtype == torch.bool and not bool(masks.any())  # the evaluator passes an all-False padding mask
        base = inputs[:, -1:]
        if text_embeddings is not None:
            base = base + 0.1 * text_embeddings.mean((1, 2))[:, None]
        return base.expand(-1, horizon).contiguous()


def _batches(with_text):
    g = torch.Generator().manual_seed(3)
    out = []
    for n in (5, 3, 7):  # ragged batch sizes: the mean must be sample weighted
        b = {"context": torch.randn(n, 64, generator=g), "horizon": torch.randn(n, 16, generator=g)}
        if with_text:
            b["text_embeddings"] = torch.randn(n, 2, 8, generator=g)
        out.append(b)
    return out


@pytest.mark.parametrize("with_text", [False, True])
def test_metrics_are_sample_weighted_and_match_reference(with_text):
    batches = _batches(with_text)
    got = MultimodalEvaluator(_Stub(), torch.device("cpu")).evaluate(batches)
    se = ae = n = 0.0
    for b in batches:
        p = _Stub()(16, b["context"], torch.zeros_like(b["context"], dtype=torch.bool), b.get("text_embeddings"))
        se += ((p - b["horizon"]) ** 2).mean().item() * len(p)
        ae += (p - b["horizon"]).abs().mean().item() * len(p)
        n += len(p)
    assert got["mse"] == pytest.approx(se / n, rel=1e-6) and got["mae"] == pytest.approx(ae / n, rel=1e-6)
    if REF.exists():
        sys.path.insert(0, str(REF))
        try:
            from tsfmx.evaluator import MultimodalEvaluator as RefEvaluator
        finally:
            sys.path.remove(str(REF))
        ref = RefEvaluator(_Stub(), torch.device("cpu")).evaluate(batches)
        assert got["mse"] == pytest.approx(ref["mse"], rel=1e-6) and got["mae"] == pytest.approx(ref["mae"], rel=1e-6)


def test_empty_loader_raises():
    with pytest.raises(RuntimeError, match="empty"):
        MultimodalEvaluator(_Stub(), torch.device("cpu")).evaluate([])

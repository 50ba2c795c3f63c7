"""Text -> time-series fusion (reference tsfmx/fusion.py:8-55).

``MultimodalFusion`` projects the per-patch text embeddings with 1-3 bias-free ``Linear`` + ``ReLU``
layers (ReLU after *every* layer, Xavier-uniform init) and adds the result to the patch-token
embeddings.  The parameter container (``projection.{0,2,4}.weight``) is the reference's, so reference
checkpoints load unchanged; the arithmetic runs as tcgen05 GEMMs whose epilogue applies the ReLU and,
for the last layer, the residual add that writes the fused tokens directly.
"""

from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import ACT_RELU, DT_F32, PRECISIONS, TsfmxError
from .lanes import drain


class MultimodalFusion(nn.Module):
    """Addition-based fusion of time-series and text embeddings."""

    def __init__(
        self,
        ts_embedding_dims: int,
        text_embedding_dims: int,
        num_layers: int = 1,
        hidden_dims: list[int] = [],  # noqa: B006 - same signature as the reference
    ) -> None:
        super().__init__()
        self._validate(num_layers, hidden_dims)
        dims = [text_embedding_dims, *hidden_dims, ts_embedding_dims]
        layers: list[nn.Module] = []
        for d_in, d_out in zip(dims[:-1], dims[1:]):
            layers.append(nn.Linear(d_in, d_out, bias=False))
            layers.append(nn.ReLU())
        self.projection = nn.Sequential(*layers)
        for module in self.projection.modules():
            if isinstance(module, nn.Linear):
                nn.init.xavier_uniform_(module.weight)
        self.dims = dims
        self.precision = "bf16"
        self._packed: dict[tuple, list[torch.Tensor]] = {}

    def _validate(self, num_layers: int, hidden_dims: list[int]) -> None:
        if num_layers < 1 or num_layers > 3:
            raise ValueError(f"num_layers must be between 1 and 3, got {num_layers}")
        if len(hidden_dims) != num_layers - 1:
            raise ValueError(
                f"hidden_dims must have {num_layers - 1} elements for {num_layers} layers, got {len(hidden_dims)}"
            )

    # ------------------------------------------------------------------ weights
    def linears(self) -> list[nn.Linear]:
        return [m for m in self.projection if isinstance(m, nn.Linear)]

    def _packed_weights(self, precision: int) -> list[torch.Tensor]:
        lins = self.linears()
        key = (precision, tuple((l.weight.data_ptr(), l.weight._version) for l in lins))
        if key not in self._packed:
            self._packed.clear()
            self._packed[key] = [
                ops.cast_rows(_pad_k(l.weight.detach().float()), ops.act_dtype(precision)) for l in lins
            ]
        return self._packed[key]

    # ------------------------------------------------------------------ forward
    def forward(self, ts_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
        """``ts_embeddings + relu(W_k ... relu(W_1 text))`` (reference fusion.py:44-47)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import fusion_forward_with_grad

            return fusion_forward_with_grad(self, ts_embeddings, text_embeddings)
        return self.forward_device(ts_embeddings, text_embeddings)

    def forward_device(self, ts_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
        return drain(self.forward_device_steps(ts_embeddings, text_embeddings))

    def forward_device_steps(self, ts_embeddings: torch.Tensor, text_embeddings: torch.Tensor):
        """The fusion as a step generator (one ``yield`` per kernel launch, see ``tsfmx_b200.lanes``)."""
        if not ts_embeddings.is_cuda:
            raise TsfmxError("MultimodalFusion runs on B200 only; there is no CPU fallback")
        precision = PRECISIONS[self.precision]
        adt = ops.act_dtype(precision)
        lead = ts_embeddings.shape[:-1]
        d_out = ts_embeddings.shape[-1]
        ts2 = ts_embeddings.reshape(-1, d_out).float().contiguous()
        tx2 = text_embeddings.reshape(-1, text_embeddings.shape[-1]).float().contiguous()
        if tx2.shape[0] != ts2.shape[0]:
            raise ValueError(
                f"text_embeddings {tuple(text_embeddings.shape)} do not match ts_embeddings {tuple(ts_embeddings.shape)}"
            )
        m = ts2.shape[0]
        weights = self._packed_weights(precision)
        h = ops.cast_rows(_pad_k(tx2), adt)
        yield
        for i, w in enumerate(weights):
            n = self.dims[i + 1]
            k = _round64(self.dims[i])
            last = i == len(weights) - 1
            if last:
                out = torch.empty(m, n, dtype=torch.float32, device=ts2.device)
                ops.gemm([(h, w, k)], m, n, out, DT_F32, precision=precision, act=ACT_RELU, residual=ts2)
                yield
                return out.reshape(*lead, n)
            n_pad = _round64(n)
            nxt = ops.alloc(m, n_pad, adt, ts2.device)
            if n_pad != n:
                nxt.zero_()
            ops.gemm([(h, w, k)], m, n, nxt, adt, precision=precision, act=ACT_RELU, split_off=n_pad)
            yield
            h = nxt
        raise AssertionError("unreachable")

    def freeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = False

    def unfreeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = True


def _round64(k: int) -> int:
    return (k + 63) // 64 * 64


def _pad_k(w: torch.Tensor) -> torch.Tensor:
    """Zero-pad the K (last) dimension to a multiple of 64 — one 128-byte swizzle row of bf16."""
    k = w.shape[-1]
    kp = _round64(k)
    if kp == k:
        return w.contiguous()
    out = torch.zeros(*w.shape[:-1], kp, dtype=w.dtype, device=w.device)
    out[..., :k] = w
    return out

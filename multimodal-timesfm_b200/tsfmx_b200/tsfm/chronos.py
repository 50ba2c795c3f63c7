"""Chronos-2 adapter (reference tsfmx/tsfm/chronos.py:16-207), B200-native.

The reference wraps ``chronos.Chronos2Model`` (chronos-forecasting 2.2.2); here ``Chronos2Module`` is a plain
parameter container with the upstream state-dict key names and every stage runs through the C ABI:

  preprocess  : fused instance-norm + arcsinh + patch(16) + time-encoding kernel, then the input ResidualBlock
                (48 -> 3072 -> 768, bias, ReLU) as two tcgen05 GEMMs
  forward     : [context | REG | 64 future patches] assembled once (the future-patch embeddings are batch
                invariant: zeros + fixed time encoding, so they are computed once and broadcast); per block
                QKV GEMM -> attention kernel (RoPE, no scaling, key mask) -> out GEMM -> fused residual + norm ->
                group attention as ONE GEMM on the pre-multiplied W_o W_v (group_ids = arange(B), reference
                chronos.py:117, makes it per-series) -> fused residual + norm -> ReLU MLP (2 GEMMs)
  postprocess : output ResidualBlock on the ceil(horizon / 16) patches that are needed, sinh / scale / loc and the
                (B, horizon, 21) reorder in one epilogue kernel
"""

from __future__ import annotations

import os
from types import SimpleNamespace

import torch
from torch import nn

from .. import ops
from .._lib import ACT_RELU, ACT_RELU_GRAD, DT_BF16, DT_F32, PREC_BF16, PRECISIONS, TsfmxError
from ..fusion import _pad_k
from ..lanes import drain
from .base import PreprocessResult, TsfmAdapter

QUANTILES = [0.01, 0.05] + [round(0.1 + 0.05 * i, 2) for i in range(17)] + [0.95, 0.99]


class _ResidualBlock(nn.Module):
    def __init__(self, d_in: int, d_hidden: int, d_out: int) -> None:
        super().__init__()
        self.hidden_layer = nn.Linear(d_in, d_hidden)
        self.output_layer = nn.Linear(d_hidden, d_out)
        self.residual_layer = nn.Linear(d_in, d_out)


class _LayerNorm(nn.Module):
    def __init__(self, dims: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dims))


class _MHA(nn.Module):
    def __init__(self, d_model: int, inner: int) -> None:
        super().__init__()
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)


class _AttnLayer(nn.Module):
    def __init__(self, d_model: int, inner: int) -> None:
        super().__init__()
        self.self_attention = _MHA(d_model, inner)
        self.layer_norm = _LayerNorm(d_model)


class _MLP(nn.Module):
    def __init__(self, d_model: int, d_ff: int) -> None:
        super().__init__()
        self.wi = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)


class _FFLayer(nn.Module):
    def __init__(self, d_model: int, d_ff: int) -> None:
        super().__init__()
        self.mlp = _MLP(d_model, d_ff)
        self.layer_norm = _LayerNorm(d_model)


class _Block(nn.Module):
    def __init__(self, d_model: int, inner: int, d_ff: int) -> None:
        super().__init__()
        # layer.0 = time self-attention, layer.1 = group self-attention, layer.2 = feed forward
        self.layer = nn.ModuleList([_AttnLayer(d_model, inner), _AttnLayer(d_model, inner), _FFLayer(d_model, d_ff)])


class _Encoder(nn.Module):
    def __init__(self, num_layers: int, d_model: int, inner: int, d_ff: int) -> None:
        super().__init__()
        self.block = nn.ModuleList(_Block(d_model, inner, d_ff) for _ in range(num_layers))
        self.final_layer_norm = _LayerNorm(d_model)


class Chronos2Module(nn.Module):
    """Parameters of Chronos-2 (amazon/chronos-2 layout) with upstream key names."""

    def __init__(self, num_layers: int = 12) -> None:
        super().__init__()
        self.model_dim, self.num_heads, self.d_kv, self.d_ff = 768, 12, 64, 3072
        self.eps = 1e-6
        self.rope_theta = 10000.0
        self.chronos_config = SimpleNamespace(
            context_length=8192, input_patch_size=16, input_patch_stride=16, output_patch_size=16,
            quantiles=list(QUANTILES), use_reg_token=True, use_arcsinh=True, max_output_patches=64,
            time_encoding_scale=8192,
        )
        self.config = SimpleNamespace(reg_token_id=1, pad_token_id=0, vocab_size=2)
        self.num_quantiles = len(QUANTILES)
        inner = self.num_heads * self.d_kv
        self.shared = nn.Embedding(2, self.model_dim)
        self.input_patch_embedding = _ResidualBlock(3 * 16, self.d_ff, self.model_dim)
        self.encoder = _Encoder(num_layers, self.model_dim, inner, self.d_ff)
        self.output_patch_embedding = _ResidualBlock(self.model_dim, self.d_ff, self.num_quantiles * 16)


def init_random_(model: nn.Module, seed: int = 0) -> None:
    """Deterministic init of every tensor: matrices ~ N(0, 0.03), biases ~ N(0, 0.01), norm weights 1 + 0.1 N."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "layer_norm.weight" in name:
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=gen)
            elif name.endswith(".bias"):
                v = 0.01 * torch.randn(p.shape, generator=gen)
            else:
                v = 0.03 * torch.randn(p.shape, generator=gen)
            p.copy_(v.to(p.device))


def _round64(k: int) -> int:
    return (k + 63) // 64 * 64


def _hi(t: torch.Tensor, cols: int) -> torch.Tensor:
    """relu'(h) only needs the sign: the hi half of a split activation is enough."""
    return t[:, :cols] if t.shape[1] != cols else t


class Chronos2Adapter(TsfmAdapter):
    """Adapter for Amazon Chronos-2 (120 M encoder-only model)."""

    def __init__(self, model: Chronos2Module | None = None, precision: str = "bf16") -> None:
        super().__init__()
        self._model = model if model is not None else Chronos2Module()
        self.fused_norm = os.environ.get("TSFMX_FUSED_NORM", "0") == "1"  # True: residual + RMS LayerNorm junctions in the GEMM epilogue (measured slower, kept for A/B)
        self.set_precision(precision)
        self._packed: dict[object, dict[str, object]] = {}
        mm = self._model  # RoPE frequencies as a (non-persistent) buffer: no host-to-device copy inside _weights()
        self.register_buffer(
            "_inv_freq", 1.0 / (mm.rope_theta ** (torch.arange(0, mm.d_kv, 2, dtype=torch.int64).float() / mm.d_kv)),
            persistent=False)

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision

    @property
    def model_dims(self) -> int:
        return int(self._model.model_dim)

    @property
    def patch_len(self) -> int:
        return int(self._model.chronos_config.input_patch_size)

    @property
    def num_outputs(self) -> int:
        """Channels of the ``postprocess`` output (not part of the reference contract; used for empty batches)."""
        return int(self._model.num_quantiles)

    @property
    def point_forecast_index(self) -> int:
        return list(self._model.chronos_config.quantiles).index(0.5)

    # ------------------------------------------------------------------ packed weights
    def _weights(self) -> dict[str, object]:
        prec = PRECISIONS[self.precision]
        params = list(self._model.parameters())
        key = (prec, tuple((p.data_ptr(), p._version) for p in params))
        if key in self._packed:
            return self._packed[key]
        self._packed.clear()
        m = self._model
        adt = ops.act_dtype(prec)
        dev = params[0].device

        def pack(wt: torch.Tensor) -> torch.Tensor:
            return ops.cast_rows(_pad_k(wt.detach().float()), adt)

        def f32(t: torch.Tensor) -> torch.Tensor:
            return t.detach().float().contiguous()

        ipe, ope = m.input_patch_embedding, m.output_patch_embedding
        w: dict[str, object] = {
            "in_hidden": pack(ipe.hidden_layer.weight), "in_hidden_b": f32(ipe.hidden_layer.bias),
            "in_out": pack(ipe.output_layer.weight), "in_res": pack(ipe.residual_layer.weight),
            "in_out_b": f32(ipe.output_layer.bias + ipe.residual_layer.bias),
            "out_hidden": pack(ope.hidden_layer.weight), "out_hidden_b": f32(ope.hidden_layer.bias),
            "out_out": pack(ope.output_layer.weight), "out_res": pack(ope.residual_layer.weight),
            "out_out_b": f32(ope.output_layer.bias + ope.residual_layer.bias),
            "final_ln": f32(m.encoder.final_layer_norm.weight),
            "reg": f32(m.shared.weight[m.config.reg_token_id]),
            "inv_freq": self._inv_freq.to(dev),
            "blocks": [],
        }
        for blk in m.encoder.block:
            t_att, g_att, ff = blk.layer[0], blk.layer[1], blk.layer[2]
            ta, ga = t_att.self_attention, g_att.self_attention
            # group attention with group_ids = arange(B): h + W_o W_v LN(h)  (product formed in fp64)
            w_ov = (ga.o.weight.detach().double() @ ga.v.weight.detach().double()).float()
            w["blocks"].append(
                {
                    "ln_t": f32(t_att.layer_norm.weight),
                    "qkv": pack(torch.cat([ta.q.weight, ta.k.weight, ta.v.weight], dim=0)),
                    "o": pack(ta.o.weight),
                    "ln_g": f32(g_att.layer_norm.weight),
                    "ov": pack(w_ov),
                    "ln_f": f32(ff.layer_norm.weight),
                    "wi": pack(ff.mlp.wi.weight),
                    "wo": pack(ff.mlp.wo.weight),
                }
            )
        self._packed[key] = w
        return w

    def _embed_patches(self, patched: torch.Tensor, rows: int, w: dict, prec: int, keep: dict | None = None) -> torch.Tensor:
        """input_patch_embedding: ResidualBlock 48(->64 padded) -> 3072 -> 768, bias, ReLU -> fp32 [rows, 768].
        ``keep`` (full fine-tuning) receives the block's operands for ``_embed_backward``."""
        m = self._model
        adt = ops.act_dtype(prec)
        hidden = ops.alloc(rows, m.d_ff, adt, patched.device)
        ops.gemm([(patched, w["in_hidden"], 64)], rows, m.d_ff, hidden, adt, precision=prec, act=ACT_RELU,
                 bias=w["in_hidden_b"])
        emb = torch.empty(rows, m.model_dim, dtype=torch.float32, device=patched.device)
        ops.gemm([(hidden, w["in_out"], m.d_ff), (patched, w["in_res"], 64)], rows, m.model_dim, emb, DT_F32,
                 precision=prec, bias=w["in_out_b"])
        if keep is not None:
            keep.update({"patched": patched, "hidden": hidden, "rows": rows})
        return emb

    def _future_patches(self, prec: int, dev: torch.device) -> torch.Tensor:
        """Tokenizer input rows of the 64 future patches: fixed time encoding, zero values, zero mask."""
        cc = self._model.chronos_config
        nop, ps = cc.max_output_patches, cc.output_patch_size
        x = torch.zeros(nop, 64, dtype=torch.float32, device=dev)
        x[:, :ps] = (torch.arange(0, nop * ps, dtype=torch.float32, device=dev) / cc.time_encoding_scale).view(nop, ps)
        return ops.cast_rows(x, ops.act_dtype(prec))

    def _embed_backward(self, kept: dict, d_emb: torch.Tensor, pg: dict[str, torch.Tensor]) -> None:
        """Weight / bias gradients of input_patch_embedding from dL/d(embeddings) [rows, 768] fp32, ACCUMULATED into
        ``pg`` (the block embeds the context patches and, once per step, the 64 future patches)."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        wt = self._weights_t()
        rows, patched, hidden = kept["rows"], kept["patched"], kept["hidden"]
        ps3 = 3 * m.chronos_config.input_patch_size
        d_a = ops.cast_rows(d_emb, adt)
        dh32 = torch.empty(rows, m.d_ff, dtype=torch.float32, device=d_emb.device)
        ops.gemm([(d_a, wt["in_out"], m.model_dim)], rows, m.d_ff, dh32, DT_F32, precision=prec, act=ACT_RELU_GRAD,
                 aux=_hi(hidden, m.d_ff))
        bias = ops.colsum_wgrad(d_emb)
        new = {
            "input_patch_embedding.output_layer.weight": ops.wgrad(d_a, hidden, rows, m.model_dim, m.d_ff, prec),
            "input_patch_embedding.residual_layer.weight": ops.wgrad(d_a, patched, rows, m.model_dim, 64, prec)[:, :ps3],
            "input_patch_embedding.output_layer.bias": bias,
            "input_patch_embedding.residual_layer.bias": bias,
            "input_patch_embedding.hidden_layer.weight": ops.wgrad(dh32, patched, rows, m.d_ff, 64, prec)[:, :ps3],
            "input_patch_embedding.hidden_layer.bias": ops.colsum_wgrad(dh32),
        }
        for k, v in new.items():
            pg[k] = pg[k] + v if k in pg else v.contiguous().clone()

    # ------------------------------------------------------------------ stages
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult:
        """Normalise, patch, time-encode and embed (reference chronos.py:35-60).

        masks: True = padded.  Returns embeddings (B, ceil(C/16), 768), patch masks (True = no observed point in
        the patch) and ``loc`` / ``scale`` of shape (B, 1)."""
        if not inputs.is_cuda:
            raise TsfmxError("Chronos2Adapter runs on B200 only; there is no CPU fallback")
        m, cc = self._model, self._model.chronos_config
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        if inputs.shape[-1] > cc.context_length:  # upstream truncates to the last context_length steps
            inputs, masks = inputs[..., -cc.context_length :], masks[..., -cc.context_length :]
        b = inputs.shape[0]
        patched, attn_mask, loc, scale = ops.chronos2_patchify_norm(
            inputs, masks.bool(), cc.input_patch_size, cc.use_arcsinh, float(cc.time_encoding_scale), adt, out_cols=64
        )
        n = attn_mask.shape[1]
        emb = self._embed_patches(patched, b * n, w, prec)
        return PreprocessResult(
            input_embeddings=emb.view(b, n, m.model_dim),
            masks=~attn_mask,
            normalization_stats={"loc": loc.view(b, 1), "scale": scale.view(b, 1)},
        )

    def _future_embeds(self, w: dict, prec: int, dev: torch.device) -> torch.Tensor:
        """Embeddings of the 64 all-zero future patches: only the fixed time encoding is non-zero, so they are the
        same for every series (reference chronos.py:82-100 recomputes them B times)."""
        key = ("future", prec)
        if key not in w:
            w[key] = self._embed_patches(self._future_patches(prec, dev), self._model.chronos_config.max_output_patches,
                                         w, prec)
        return w[key]

    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """Run the encoder on [context | REG | future] and return the 64 forecast positions
        (reference chronos.py:62-126) -> (batch, 64, 768)."""
        return drain(self.forward_steps(input_embeddings, masks))

    def forward_steps(self, input_embeddings: torch.Tensor, masks: torch.Tensor):
        """``forward`` as a step generator (one ``yield`` per kernel launch, see ``tsfmx_b200.lanes``)."""
        if not input_embeddings.is_cuda:
            raise TsfmxError("Chronos2Adapter runs on B200 only; there is no CPU fallback")
        m, cc = self._model, self._model.chronos_config
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        dev = input_embeddings.device
        b, n, d = input_embeddings.shape
        nop = cc.max_output_patches
        extra = 1 if cc.use_reg_token else 0
        t = n + extra + nop
        h = torch.empty(b, t, d, dtype=torch.float32, device=dev)
        h[:, :n] = input_embeddings
        if extra:
            h[:, n] = w["reg"]
        h[:, n + extra :] = self._future_embeds(w, prec, dev)
        key_mask = torch.ones(b, t, dtype=torch.bool, device=dev)
        key_mask[:, :n] = ~masks.bool()
        rows = b * t
        h2 = h.view(rows, d)
        blocks = w["blocks"]
        inner = m.num_heads * m.d_kv
        if not blocks:
            out = ops.rmsnorm(h2, w["final_ln"], m.eps, DT_F32)
            return out.view(b, t, d)[:, -nop:].contiguous()
        xn = ops.rmsnorm(h2, blocks[0]["ln_t"], m.eps, adt)
        yield
        qkv = ops.alloc(rows, 3 * inner, mid_dt, dev)
        attn = ops.alloc(rows, inner, adt, dev)
        a = ops.alloc(rows, d, mid_dt, dev)
        u = ops.alloc(rows, m.d_ff, adt, dev)
        final = None
        for i, bw in enumerate(blocks):
            ops.gemm([(xn, bw["qkv"], d)], rows, 3 * inner, qkv, mid_dt, precision=prec)
            yield
            ops.encoder_attention(qkv, b, t, m.num_heads, m.d_kv, key_mask, w["inv_freq"], adt, out=attn)
            yield
            has_next = i + 1 < len(blocks)
            if not has_next:
                final = torch.empty(rows, d, dtype=torch.float32, device=dev)
            if self.fused_norm:
                # residual add + next RMS LayerNorm in the GEMM epilogue (3-CTA clusters own a 768-wide row panel)
                ops.gemm_rownorm(attn, bw["o"], inner, rows, d, prec, None, bw["ln_g"], h2, h2, adt, xn, m.eps)
                yield
                ops.gemm_rownorm(xn, bw["ov"], d, rows, d, prec, None, bw["ln_f"], h2, h2, adt, xn, m.eps)
                yield
                ops.gemm([(xn, bw["wi"], d)], rows, m.d_ff, u, adt, precision=prec, act=ACT_RELU)
                yield
                if has_next:
                    ops.gemm_rownorm(u, bw["wo"], m.d_ff, rows, d, prec, None, blocks[i + 1]["ln_t"], h2, h2, adt, xn, m.eps)
                    yield
                else:
                    ops.gemm_rownorm(u, bw["wo"], m.d_ff, rows, d, prec, None, w["final_ln"], h2, h2, DT_F32, final, m.eps)
                    yield
            else:
                ops.gemm([(attn, bw["o"], inner)], rows, d, a, mid_dt, precision=prec)
                yield
                ops.norm_residual_norm(a, h2, None, bw["ln_g"], m.eps, h2, adt, xn)
                yield
                ops.gemm([(xn, bw["ov"], d)], rows, d, a, mid_dt, precision=prec)
                yield
                ops.norm_residual_norm(a, h2, None, bw["ln_f"], m.eps, h2, adt, xn)
                yield
                ops.gemm([(xn, bw["wi"], d)], rows, m.d_ff, u, adt, precision=prec, act=ACT_RELU)
                yield
                ops.gemm([(u, bw["wo"], m.d_ff)], rows, d, a, mid_dt, precision=prec)
                yield
                if has_next:
                    ops.norm_residual_norm(a, h2, None, blocks[i + 1]["ln_t"], m.eps, h2, adt, xn)
                    yield
                else:
                    ops.norm_residual_norm(a, h2, None, w["final_ln"], m.eps, h2, DT_F32, final)
                    yield
        return final.view(b, t, d)[:, -nop:].contiguous()

    def postprocess(
        self,
        horizon: int,
        output_embeddings: torch.Tensor,
        normalization_stats: dict[str, torch.Tensor],
    ) -> torch.Tensor:
        """Quantile head + inverse instance norm (reference chronos.py:128-169) -> (batch, horizon, 21).

        Raises ValueError if ``horizon`` exceeds 64 patches * 16 steps."""
        m, cc = self._model, self._model.chronos_config
        nop, ps = cc.max_output_patches, cc.output_patch_size
        max_horizon = nop * ps
        if horizon > max_horizon:
            raise ValueError(
                f"horizon ({horizon}) exceeds the maximum prediction length "
                f"({max_horizon} = {nop} patches * {ps} steps)."
            )
        if not output_embeddings.is_cuda:
            raise TsfmxError("Chronos2Adapter runs on B200 only; there is no CPU fallback")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        b, _, d = output_embeddings.shape
        used = (horizon + ps - 1) // ps  # only these output patches reach the forecast
        x = output_embeddings[:, :used].float().reshape(b * used, d).contiguous()
        xa = ops.cast_rows(x, adt)
        rows = b * used
        hid = ops.alloc(rows, m.d_ff, adt, x.device)
        ops.gemm([(xa, w["out_hidden"], d)], rows, m.d_ff, hid, adt, precision=prec, act=ACT_RELU, bias=w["out_hidden_b"])
        nq = m.num_quantiles
        preds = torch.empty(rows, nq * ps, dtype=torch.float32, device=x.device)
        ops.gemm([(hid, w["out_out"], m.d_ff), (xa, w["out_res"], d)], rows, nq * ps, preds, DT_F32, precision=prec,
                 bias=w["out_out_b"])
        return ops.chronos2_finalize(preds, b, used, nq, ps, horizon, cc.use_arcsinh, normalization_stats["loc"],
                                     normalization_stats["scale"])

    # ------------------------------------------------------------------ training path (frozen backbone)
    def _weights_t(self) -> dict[str, object]:
        """Transposed (dgrad) packs of the frozen weights: dX = dY W as a K-major GEMM on W^T."""
        w = self._weights()
        if "t" not in w:
            prec = PRECISIONS[self.precision]
            adt = ops.act_dtype(prec)
            m = self._model

            def pack_t(wt: torch.Tensor) -> torch.Tensor:
                return ops.cast_rows(_pad_k(wt.detach().float().t().contiguous()), adt)

            ope = m.output_patch_embedding
            blocks = []
            for blk in m.encoder.block:
                t_att, g_att, ff = blk.layer[0], blk.layer[1], blk.layer[2]
                ta, ga = t_att.self_attention, g_att.self_attention
                w_ov = (ga.o.weight.detach().double() @ ga.v.weight.detach().double()).float()
                blocks.append({
                    "qkv": pack_t(torch.cat([ta.q.weight, ta.k.weight, ta.v.weight], dim=0)), "o": pack_t(ta.o.weight),
                    "ov": pack_t(w_ov), "wi": pack_t(ff.mlp.wi.weight), "wo": pack_t(ff.mlp.wo.weight),
                })
            w["t"] = {"blocks": blocks, "out_hidden": pack_t(ope.hidden_layer.weight),
                      "out_out": pack_t(ope.output_layer.weight), "out_res": pack_t(ope.residual_layer.weight),
                      "in_out": pack_t(m.input_patch_embedding.output_layer.weight)}
        return w["t"]

    def forward_saving(self, input_embeddings: torch.Tensor, masks: torch.Tensor, for_wgrad: bool = False):
        """``forward`` that keeps what the activation-gradient pass needs: the residual stream at the input of every
        RMS LayerNorm, the raw qkv of every time attention and the ReLU outputs of every MLP.  ``for_wgrad`` (full
        fine-tuning) also keeps the input of every GEMM and the operands of the future-patch embedding."""
        m, cc = self._model, self._model.chronos_config
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        dev = input_embeddings.device
        b, n, d = input_embeddings.shape
        nop = cc.max_output_patches
        extra = 1 if cc.use_reg_token else 0
        t = n + extra + nop
        h = torch.empty(b, t, d, dtype=torch.float32, device=dev)
        h[:, :n] = input_embeddings
        if extra:
            h[:, n] = w["reg"]
        future_kept: dict | None = {} if for_wgrad else None
        if for_wgrad:  # the future-patch embeddings depend on the weights being trained: fresh, with operands kept
            h[:, n + extra:] = self._embed_patches(self._future_patches(prec, dev), nop, w, prec, keep=future_kept)
        else:
            h[:, n + extra:] = self._future_embeds(w, prec, dev)
        key_mask = torch.ones(b, t, dtype=torch.bool, device=dev)
        key_mask[:, :n] = ~masks.bool()
        rows = b * t
        cur = h.view(rows, d)
        blocks = w["blocks"]
        inner = m.num_heads * m.d_kv
        saved = {"shape": (b, n, t, d), "key_mask": key_mask, "blocks": [], "future": future_kept}
        xn = ops.rmsnorm(cur, blocks[0]["ln_t"], m.eps, adt) if blocks else None
        attn = ops.alloc(rows, inner, adt, dev)
        a = ops.alloc(rows, d, mid_dt, dev)
        for i, bw in enumerate(blocks):
            qkv = ops.alloc(rows, 3 * inner, mid_dt, dev)
            u = ops.alloc(rows, m.d_ff, adt, dev)
            h1 = torch.empty(rows, d, dtype=torch.float32, device=dev)
            h2 = torch.empty(rows, d, dtype=torch.float32, device=dev)
            h3 = torch.empty(rows, d, dtype=torch.float32, device=dev)
            xn_t = xn  # normed block input (operand of the qkv GEMM)
            if for_wgrad:  # the weight-gradient GEMMs read these later: one set per block instead of reused scratch
                attn = ops.alloc(rows, inner, adt, dev)
                xn_g, xn_f = ops.alloc(rows, d, adt, dev), ops.alloc(rows, d, adt, dev)
            else:
                xn_g = xn_f = xn
            ops.gemm([(xn_t, bw["qkv"], d)], rows, 3 * inner, qkv, mid_dt, precision=prec)
            ops.encoder_attention(qkv, b, t, m.num_heads, m.d_kv, key_mask, w["inv_freq"], adt, out=attn)
            ops.gemm([(attn, bw["o"], inner)], rows, d, a, mid_dt, precision=prec)
            ops.norm_residual_norm(a, cur, None, bw["ln_g"], m.eps, h1, adt, xn_g)
            ops.gemm([(xn_g, bw["ov"], d)], rows, d, a, mid_dt, precision=prec)
            ops.norm_residual_norm(a, h1, None, bw["ln_f"], m.eps, h2, adt, xn_f)
            ops.gemm([(xn_f, bw["wi"], d)], rows, m.d_ff, u, adt, precision=prec, act=ACT_RELU)
            ops.gemm([(u, bw["wo"], m.d_ff)], rows, d, a, mid_dt, precision=prec)
            nxt = blocks[i + 1]["ln_t"] if i + 1 < len(blocks) else w["final_ln"]
            last = i + 1 == len(blocks)
            final = torch.empty(rows, d, dtype=torch.float32, device=dev) if last else None
            xn = None if last else (ops.alloc(rows, d, adt, dev) if for_wgrad else xn)
            ops.norm_residual_norm(a, h2, None, nxt, m.eps, h3, DT_F32 if last else adt, final if last else xn)
            entry = {"h0": cur, "h1": h1, "h2": h2, "qkv": qkv, "u": u}
            if for_wgrad:
                entry.update({"xn_t": xn_t, "attn": attn, "xn_g": xn_g, "xn_f": xn_f})
            saved["blocks"].append(entry)
            cur = h3
        if not blocks:
            final = ops.rmsnorm(cur, w["final_ln"], m.eps, DT_F32)
        saved["h_last"] = cur
        return final.view(b, t, d)[:, -nop:].contiguous(), saved

    def forward_backward(self, saved, d_out: torch.Tensor, param_grads: dict[str, torch.Tensor] | None = None,
                         on_grads=None) -> torch.Tensor:
        """dL/d(returned forecast-position embeddings) [B * 64, D] -> dL/d(input embeddings) [B * N, D] fp32.

        ``param_grads`` (full fine-tuning, needs ``forward_saving(for_wgrad=True)``): receives the gradient of every
        encoder parameter under its state-dict name, of the [REG] embedding, and the future patches' share of the
        input_patch_embedding gradients.  ``on_grads(tensors)``: called with each finished block's gradients."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w, wt = self._weights(), self._weights_t()
        b, n, t, d = saved["shape"]
        rows = b * t
        dev = d_out.device
        nop = m.chronos_config.max_output_patches
        inner = m.num_heads * m.d_kv
        g_final = torch.zeros(b, t, d, dtype=torch.float32, device=dev)
        g_final[:, -nop:] = d_out.view(b, nop, d)
        g = torch.empty(rows, d, dtype=torch.float32, device=dev)  # running dL/d(residual stream)
        ops.rmsnorm_bwd_chain(None, saved["h_last"], w["final_ln"], g_final.view(rows, d), None, None, m.eps, g, adt, None,
                              rows, d)
        pg = param_grads
        if pg is not None:
            pg["encoder.final_layer_norm.weight"] = ops.colsum_wgrad(g_final.view(rows, d), saved["h_last"], m.eps)
        du = ops.alloc(rows, m.d_ff, adt, dev)
        dxn = ops.alloc(rows, d, mid_dt, dev)
        dattn = ops.alloc(rows, inner, mid_dt, dev)
        dqkv = ops.alloc(rows, 3 * inner, adt, dev)
        blocks, tblocks = w["blocks"], wt["blocks"]
        for i in reversed(range(len(blocks))):
            bw, tw, s = blocks[i], tblocks[i], saved["blocks"][i]
            pre = f"encoder.block.{i}.layer."
            if pg is not None and on_grads is not None and i + 1 < len(blocks):
                done = f"encoder.block.{i + 1}."  # the block above is complete: hand its gradients over
                on_grads([v for k, v in pg.items() if k.startswith(done)])
            # feed-forward: h3 = h2 + relu(rms(h2) Wi) Wo
            ga = ops.cast_rows(g, adt)
            if pg is not None:
                pg[pre + "2.mlp.wo.weight"] = ops.wgrad(ga, s["u"], rows, d, m.d_ff, prec)
            ops.gemm([(ga, tw["wo"], d)], rows, m.d_ff, du, adt, precision=prec, act=ACT_RELU_GRAD, aux=_hi(s["u"], m.d_ff),
                     split_off=m.d_ff)
            ops.gemm([(du, tw["wi"], m.d_ff)], rows, d, dxn, mid_dt, precision=prec)
            if pg is not None:
                pg[pre + "2.mlp.wi.weight"] = ops.wgrad(du, s["xn_f"], rows, m.d_ff, d, prec)
                pg[pre + "2.layer_norm.weight"] = ops.colsum_wgrad(dxn, s["h2"], m.eps)
            ops.rmsnorm_bwd_chain(g, s["h2"], bw["ln_f"], dxn, None, None, m.eps, g, adt, None, rows, d)
            # group attention (degenerate: one GEMM): h2 = h1 + rms(h1) Wov
            ga = ops.cast_rows(g, adt)
            ops.gemm([(ga, tw["ov"], d)], rows, d, dxn, mid_dt, precision=prec)
            if pg is not None:
                # W_ov = W_o W_v: dW_o = dW_ov W_v^T, dW_v = W_o^T dW_ov (two 768^3 products in parameter space); with
                # group_ids = arange(B) every softmax of the group attention has ONE key, so q and k get no gradient
                att = m.encoder.block[i].layer[1].self_attention
                d_ov = ops.wgrad(ga, s["xn_g"], rows, d, d, prec)
                pg[pre + "1.self_attention.o.weight"] = d_ov @ att.v.weight.detach().float().t()
                pg[pre + "1.self_attention.v.weight"] = att.o.weight.detach().float().t() @ d_ov
                pg[pre + "1.self_attention.q.weight"] = torch.zeros_like(att.q.weight, dtype=torch.float32)
                pg[pre + "1.self_attention.k.weight"] = torch.zeros_like(att.k.weight, dtype=torch.float32)
                pg[pre + "1.layer_norm.weight"] = ops.colsum_wgrad(dxn, s["h1"], m.eps)
            ops.rmsnorm_bwd_chain(g, s["h1"], bw["ln_g"], dxn, None, None, m.eps, g, adt, None, rows, d)
            # time attention: h1 = h0 + attn(rms(h0) Wqkv) Wo
            ga = ops.cast_rows(g, adt)
            if pg is not None:
                pg[pre + "0.self_attention.o.weight"] = ops.wgrad(ga, s["attn"], rows, d, inner, prec)
            ops.gemm([(ga, tw["o"], d)], rows, inner, dattn, mid_dt, precision=prec)
            ops.encoder_attention_bwd(s["qkv"], dattn, b, t, m.num_heads, m.d_kv, saved["key_mask"], w["inv_freq"], adt,
                                      dqkv=dqkv)
            ops.gemm([(dqkv, tw["qkv"], 3 * inner)], rows, d, dxn, mid_dt, precision=prec)
            if pg is not None:
                d_qkv_w = ops.wgrad(dqkv, s["xn_t"], rows, 3 * inner, d, prec)
                for j, name in enumerate("qkv"):
                    pg[pre + f"0.self_attention.{name}.weight"] = d_qkv_w[j * inner : (j + 1) * inner]
                pg[pre + "0.layer_norm.weight"] = ops.colsum_wgrad(dxn, s["h0"], m.eps)
            ops.rmsnorm_bwd_chain(g, s["h0"], bw["ln_t"], dxn, None, None, m.eps, g, adt, None, rows, d)
        g3 = g.view(b, t, d)
        if pg is not None:
            if on_grads is not None and blocks:
                on_grads([v for k, v in pg.items() if k.startswith("encoder.block.0.")])
            cc = m.chronos_config
            extra = 1 if cc.use_reg_token else 0
            reg = torch.zeros_like(m.shared.weight, dtype=torch.float32)
            if extra:
                reg[m.config.reg_token_id] = g3[:, n].sum(0)  # the [REG] embedding row feeds position n of every series
            pg["shared.weight"] = reg
            # the 64 future-patch embeddings are shared by every series: their gradient is the batch sum
            self._embed_backward(saved["future"], g3[:, n + extra:].sum(0).contiguous(), pg)
        return g3[:, :n].reshape(b * n, d).contiguous()

    def preprocess_saving(self, inputs: torch.Tensor, masks: torch.Tensor):
        """``preprocess`` that keeps the input ResidualBlock's operands (full fine-tuning)."""
        if not inputs.is_cuda:
            raise TsfmxError("Chronos2Adapter runs on B200 only; there is no CPU fallback")
        m, cc = self._model, self._model.chronos_config
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        if inputs.shape[-1] > cc.context_length:
            inputs, masks = inputs[..., -cc.context_length:], masks[..., -cc.context_length:]
        b = inputs.shape[0]
        patched, attn_mask, loc, scale = ops.chronos2_patchify_norm(
            inputs, masks.bool(), cc.input_patch_size, cc.use_arcsinh, float(cc.time_encoding_scale), adt, out_cols=64
        )
        n = attn_mask.shape[1]
        kept: dict = {}
        emb = self._embed_patches(patched, b * n, w, prec, keep=kept)
        pre = PreprocessResult(emb.view(b, n, m.model_dim), ~attn_mask, {"loc": loc.view(b, 1), "scale": scale.view(b, 1)})
        return pre, kept

    def preprocess_backward(self, saved, d_emb: torch.Tensor, param_grads: dict[str, torch.Tensor]) -> None:
        """Context patches' share of the input_patch_embedding gradients (added to the future patches' share that
        ``forward_backward`` left in ``param_grads``)."""
        self._embed_backward(saved, d_emb, param_grads)

    def postprocess_saving(self, horizon: int, output_embeddings: torch.Tensor, normalization_stats):
        """``postprocess`` that keeps the head's ReLU output and raw predictions for the backward pass."""
        m, cc = self._model, self._model.chronos_config
        nop, ps = cc.max_output_patches, cc.output_patch_size
        if horizon > nop * ps:
            raise ValueError(
                f"horizon ({horizon}) exceeds the maximum prediction length ({nop * ps} = {nop} patches * {ps} steps)."
            )
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        b, _, d = output_embeddings.shape
        used = (horizon + ps - 1) // ps
        x = output_embeddings[:, :used].float().reshape(b * used, d).contiguous()
        xa = ops.cast_rows(x, adt)
        rows = b * used
        hid = ops.alloc(rows, m.d_ff, adt, x.device)
        ops.gemm([(xa, w["out_hidden"], d)], rows, m.d_ff, hid, adt, precision=prec, act=ACT_RELU, bias=w["out_hidden_b"])
        nq = m.num_quantiles
        preds = torch.empty(rows, nq * ps, dtype=torch.float32, device=x.device)
        ops.gemm([(hid, w["out_out"], m.d_ff), (xa, w["out_res"], d)], rows, nq * ps, preds, DT_F32, precision=prec,
                 bias=w["out_out_b"])
        out = ops.chronos2_finalize(preds, b, used, nq, ps, horizon, cc.use_arcsinh, normalization_stats["loc"],
                                    normalization_stats["scale"])
        saved = {"hid": hid, "preds": preds, "scale": normalization_stats["scale"].reshape(b), "b": b, "used": used,
                 "horizon": horizon, "d": d, "xa": xa}
        return out, saved

    def postprocess_backward(self, saved, grad_forecast: torch.Tensor,
                             param_grads: dict[str, torch.Tensor] | None = None) -> torch.Tensor:
        """dL/d(forecast) [B, h, 21] -> dL/d(output embeddings) [B, 64, D] fp32 (non-zero in the used patches);
        ``param_grads`` additionally receives the output ResidualBlock's gradients (full fine-tuning)."""
        m, cc = self._model, self._model.chronos_config
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        wt = self._weights_t()
        b, used, horizon, d = saved["b"], saved["used"], saved["horizon"], saved["d"]
        nop, ps, nq = cc.max_output_patches, cc.output_patch_size, m.num_quantiles
        rows = b * used
        dev = grad_forecast.device
        # undo the finalize epilogue: out[b, t, q] = sinh(preds[b, t // ps, q, t % ps]) * scale[b] + loc[b]
        g = torch.zeros(b, used * ps, nq, dtype=torch.float32, device=dev)
        g[:, :horizon] = grad_forecast.reshape(b, horizon, nq) * saved["scale"][:, None, None]
        g = g.view(b, used, ps, nq).permute(0, 1, 3, 2).reshape(rows, nq * ps)
        if cc.use_arcsinh:
            g = g * torch.cosh(saved["preds"])
        kp = _round64(nq * ps)
        dpre32 = torch.zeros(rows, kp, dtype=torch.float32, device=dev)
        dpre32[:, : nq * ps] = g
        dpre = ops.cast_rows(dpre32, adt)
        if param_grads is None:
            dhid = ops.alloc(rows, m.d_ff, adt, dev)
            ops.gemm([(dpre, wt["out_out"], kp)], rows, m.d_ff, dhid, adt, precision=prec, act=ACT_RELU_GRAD,
                     aux=_hi(saved["hid"], m.d_ff), split_off=m.d_ff)
        else:
            dhid32 = torch.empty(rows, m.d_ff, dtype=torch.float32, device=dev)
            ops.gemm([(dpre, wt["out_out"], kp)], rows, m.d_ff, dhid32, DT_F32, precision=prec, act=ACT_RELU_GRAD,
                     aux=_hi(saved["hid"], m.d_ff))
            dhid = ops.cast_rows(dhid32, adt)
            head, width = "output_patch_embedding.", nq * ps
            bias = ops.colsum_wgrad(dpre32)[:width].contiguous()
            param_grads[head + "output_layer.weight"] = ops.wgrad(dpre, saved["hid"], rows, kp, m.d_ff, prec)[:width]
            param_grads[head + "residual_layer.weight"] = ops.wgrad(dpre, saved["xa"], rows, kp, d, prec)[:width]
            param_grads[head + "output_layer.bias"] = bias
            param_grads[head + "residual_layer.bias"] = bias.clone()
            param_grads[head + "hidden_layer.weight"] = ops.wgrad(dhid32, saved["xa"], rows, m.d_ff, d, prec)
            param_grads[head + "hidden_layer.bias"] = ops.colsum_wgrad(dhid32)
        dx = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ops.gemm([(dhid, wt["out_hidden"], m.d_ff), (dpre, wt["out_res"], kp)], rows, d, dx, DT_F32, precision=prec)
        d_out = torch.zeros(b, nop, d, dtype=torch.float32, device=dev)
        d_out[:, :used] = dx.view(b, used, d)
        return d_out

    # ------------------------------------------------------------------ checkpoints / freezing
    def load_checkpoint(self, path: str) -> None:
        """Load an upstream Chronos-2 safetensors checkpoint (strict), reference chronos.py:171-174."""
        from safetensors.torch import load_file

        self._model.load_state_dict(load_file(path), strict=True)

    @classmethod
    def from_pretrained(cls, device: torch.device, repo_id: str = "amazon/chronos-2") -> "Chronos2Adapter":
        """Download + load pretrained weights (reference chronos.py:176-199); needs network access."""
        from huggingface_hub import hf_hub_download

        instance = cls(Chronos2Module())
        instance.to(device)
        instance.load_checkpoint(hf_hub_download(repo_id=repo_id, filename="model.safetensors"))
        return instance

    @property
    def graph_safe(self) -> bool:
        """The forecast path makes no host synchronisation (lazily built tables are filled by the eager pass that
        precedes a capture), so ``MultimodalDecoder`` / ``MultimodalEvaluator`` may replay it from a CUDA graph."""
        return True

    def freeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = False

    def unfreeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = True

"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/timesfm_oracle.py header for the import rules).

Chronos-2 path of the reference: ``Chronos2Adapter`` (/root/reference/src/tsfmx/tsfm/chronos.py:35-169) wraps
``chronos.Chronos2Model`` of ``chronos-forecasting`` 2.2.2 (/root/reference/uv.lock:131-133), which is neither
vendored nor installable here.  This file therefore

* restates the published Chronos-2 architecture (amazon/chronos-2 configuration:
  d_model 768, d_kv 64, 12 heads, d_ff 3072, 12 blocks, ReLU, T5-style RMS LayerNorm eps 1e-6, RoPE theta 10000,
  input/output patch 16, 21 quantiles, [REG] token, arcsinh instance norm, 64 output patches, time-encoding scale
  8192) in plain fp32 torch, INCLUDING the O(B^2) group attention exactly as upstream runs it, and
* follows the reference adapter's control flow line by line around it.

PARITY PINNING: **parity unpinned** against upstream — there is no upstream source, test or golden vector to pin this
restatement to.  What CAN be pinned is (tests/test_oracle_cpu.py): the sub-blocks that coincide with
``transformers.models.t5`` (RMS LayerNorm == T5LayerNorm, the ReLU feed-forward sub-layer == T5LayerFF, the bias-free
un-scaled attention of the group sub-layer == T5Attention without relative bias) and the rotary embedding
(== transformers' Llama ``apply_rotary_pos_emb``, same inv_freq), on identical weights; plus self-consistency (shapes,
mask conventions, the degenerate-group-attention identity, the inverse-norm round trip).  Still recalled, unpinned:
the context preparation (instance-norm statistics, arcsinh, NaN left padding, the sign / scale of the time encoding),
the [REG] token and future-patch layout, the group-attention masking rule, the head's (quantile, patch) output
layout and the quantile list.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import nn

from .timesfm_oracle import PreprocessResult

QUANTILES = [0.01, 0.05] + [round(0.1 + 0.05 * i, 2) for i in range(17)] + [0.95, 0.99]


@dataclass
class Chronos2Config:
    d_model: int = 768
    d_kv: int = 64
    num_heads: int = 12
    d_ff: int = 3072
    num_layers: int = 12
    layer_norm_epsilon: float = 1e-6
    rope_theta: float = 10000.0
    vocab_size: int = 2
    pad_token_id: int = 0
    reg_token_id: int = 1
    context_length: int = 8192
    input_patch_size: int = 16
    input_patch_stride: int = 16
    output_patch_size: int = 16
    quantiles: list[float] = field(default_factory=lambda: list(QUANTILES))
    use_reg_token: bool = True
    use_arcsinh: bool = True
    max_output_patches: int = 64
    time_encoding_scale: int = 8192


class RMSLayerNorm(nn.Module):
    """T5-style: no mean subtraction, no bias."""

    def __init__(self, dim: int, eps: float) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps

    def forward(self, x):
        variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        return self.weight * (x * torch.rsqrt(variance + self.variance_epsilon))


class ResidualBlock(nn.Module):
    """output_layer(act(hidden_layer(x))) + residual_layer(x), with bias, ReLU."""

    def __init__(self, in_dim: int, h_dim: int, out_dim: int) -> None:
        super().__init__()
        self.hidden_layer = nn.Linear(in_dim, h_dim)
        self.output_layer = nn.Linear(h_dim, out_dim)
        self.residual_layer = nn.Linear(in_dim, out_dim)

    def forward(self, x):
        return self.output_layer(torch.relu(self.hidden_layer(x))) + self.residual_layer(x)


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


class MHA(nn.Module):
    """Bias-less q/k/v/o, NO 1/sqrt(d) scaling, optional RoPE, additive mask, fp32 softmax."""

    def __init__(self, cfg: Chronos2Config, use_rope: bool) -> None:
        super().__init__()
        self.n_heads, self.d_kv = cfg.num_heads, cfg.d_kv
        inner = cfg.num_heads * cfg.d_kv
        self.q = nn.Linear(cfg.d_model, inner, bias=False)
        self.k = nn.Linear(cfg.d_model, inner, bias=False)
        self.v = nn.Linear(cfg.d_model, inner, bias=False)
        self.o = nn.Linear(inner, cfg.d_model, bias=False)
        self.use_rope = use_rope
        if use_rope:
            inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.d_kv, 2, dtype=torch.int64).float() / cfg.d_kv))
            self.register_buffer("inv_freq", inv_freq, persistent=False)

    def forward(self, hidden, mask, position_ids=None):
        lead, seq = hidden.shape[0], hidden.shape[1]

        def shape(t):
            return t.view(lead, seq, self.n_heads, self.d_kv).transpose(1, 2)

        q, k, v = shape(self.q(hidden)), shape(self.k(hidden)), shape(self.v(hidden))
        if self.use_rope:
            freqs = position_ids[:, :, None].float() * self.inv_freq[None, None, :]
            emb = torch.cat((freqs, freqs), dim=-1)
            cos, sin = emb.cos()[:, None], emb.sin()[:, None]
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
        scores = torch.matmul(q, k.transpose(3, 2)) + mask
        weights = torch.softmax(scores.float(), dim=-1).type_as(scores)
        out = torch.matmul(weights, v).transpose(1, 2).reshape(lead, seq, self.n_heads * self.d_kv)
        return self.o(out)


class EncoderBlock(nn.Module):
    def __init__(self, cfg: Chronos2Config) -> None:
        super().__init__()
        eps = cfg.layer_norm_epsilon
        self.time_attn, self.time_ln = MHA(cfg, use_rope=True), RMSLayerNorm(cfg.d_model, eps)
        self.group_attn, self.group_ln = MHA(cfg, use_rope=False), RMSLayerNorm(cfg.d_model, eps)
        self.wi = nn.Linear(cfg.d_model, cfg.d_ff, bias=False)
        self.wo = nn.Linear(cfg.d_ff, cfg.d_model, bias=False)
        self.ff_ln = RMSLayerNorm(cfg.d_model, eps)

    def forward(self, h, position_ids, time_mask, group_time_mask):
        h = h + self.time_attn(self.time_ln(h), time_mask, position_ids)
        ht = h.transpose(0, 1)  # (T, B, D): group attention runs along the BATCH axis
        ht = ht + self.group_attn(self.group_ln(ht), group_time_mask)
        h = ht.transpose(0, 1)
        return h + self.wo(torch.relu(self.wi(self.ff_ln(h))))


class Chronos2Model(nn.Module):
    """The attributes / methods the reference adapter touches (chronos.py:25-166)."""

    def __init__(self, cfg: Chronos2Config | None = None) -> None:
        super().__init__()
        self.cfg = cfg = cfg or Chronos2Config()
        self.model_dim = cfg.d_model
        self.num_quantiles = len(cfg.quantiles)
        self.shared = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.input_patch_embedding = ResidualBlock(cfg.input_patch_size * 3, cfg.d_ff, cfg.d_model)
        self.blocks = nn.ModuleList(EncoderBlock(cfg) for _ in range(cfg.num_layers))
        self.final_layer_norm = RMSLayerNorm(cfg.d_model, cfg.layer_norm_epsilon)
        self.output_patch_embedding = ResidualBlock(cfg.d_model, cfg.d_ff, self.num_quantiles * cfg.output_patch_size)
        self.eval()

    # --- InstanceNorm (use_arcsinh) + Patch + time encoding; ignores context_mask for the statistics
    def _prepare_patched_context(self, context, context_mask):
        cfg = self.cfg
        context_mask = context_mask.to(context.dtype)
        if context.shape[-1] > cfg.context_length:
            context, context_mask = context[..., -cfg.context_length :], context_mask[..., -cfg.context_length :]
        x = context.to(torch.float32)
        loc = torch.nan_to_num(torch.nanmean(x, dim=-1, keepdim=True), nan=0.0)
        scale = torch.nan_to_num((x - loc).square().nanmean(dim=-1, keepdim=True).sqrt(), nan=1.0)
        scale = torch.where(scale == 0, torch.full_like(scale, 1e-5), scale)
        x = (x - loc) / scale
        if cfg.use_arcsinh:
            x = torch.arcsinh(x)
        p = cfg.input_patch_size

        def patch(t):
            length = t.shape[-1]
            if length % p != 0:
                pad = torch.full((*t.shape[:-1], p - length % p), float("nan"), dtype=t.dtype)
                t = torch.cat((pad, t), dim=-1)
            return t.unfold(-1, p, cfg.input_patch_stride)

        patched_context = patch(x)
        patched_mask = torch.nan_to_num(patch(context_mask), nan=0.0)
        patched_context = torch.where(patched_mask > 0.0, patched_context, 0.0)
        attention_mask = patched_mask.sum(dim=-1) > 0
        n = attention_mask.shape[-1]
        time_enc = torch.arange(-n * p, 0, dtype=torch.float32).reshape(1, n, p).expand(context.shape[0], -1, -1)
        time_enc = time_enc.div(cfg.time_encoding_scale)
        return torch.cat([time_enc, patched_context, patched_mask], dim=-1), attention_mask, (loc, scale)

    def encoder(self, inputs_embeds, group_ids, attention_mask):
        b, t, _ = inputs_embeds.shape
        fmin = torch.finfo(inputs_embeds.dtype).min
        position_ids = torch.arange(t)[None, :]
        time_mask = (1.0 - attention_mask[:, None, None, :]) * fmin  # (B, 1, 1, T)
        group_mask = (group_ids[:, None] == group_ids[None, :]).to(inputs_embeds.dtype)
        gtm = torch.einsum("qb,bt->qbt", group_mask, attention_mask)
        gtm = (1.0 - gtm.permute(2, 0, 1)[:, None]) * fmin  # (T, 1, Q, B)
        h = inputs_embeds
        for block in self.blocks:
            h = block(h, position_ids, time_mask, gtm)
        return (self.final_layer_norm(h),)

    def instance_norm_inverse(self, x, loc_scale):
        loc, scale = loc_scale
        return torch.sinh(x) * scale + loc if self.cfg.use_arcsinh else x * scale + loc


class OracleChronos2Adapter(nn.Module):
    """Reference ``Chronos2Adapter`` control flow (chronos.py:35-169) around the restated model."""

    def __init__(self, model: Chronos2Model) -> None:
        super().__init__()
        self._model = model

    @property
    def model_dims(self) -> int:
        return self._model.model_dim

    @property
    def patch_len(self) -> int:
        return self._model.cfg.input_patch_size

    @property
    def point_forecast_index(self) -> int:
        return list(self._model.cfg.quantiles).index(0.5)

    def preprocess(self, inputs, masks):  # chronos.py:35-60
        context_mask = (~masks).to(inputs.dtype)
        patched, attention_mask, (loc, scale) = self._model._prepare_patched_context(inputs, context_mask)
        return PreprocessResult(
            input_embeddings=self._model.input_patch_embedding(patched),
            masks=attention_mask == 0,
            normalization_stats={"loc": loc, "scale": scale},
        )

    def forward(self, input_embeddings, masks):  # chronos.py:62-126
        cfg = self._model.cfg
        b, dtype = input_embeddings.shape[0], input_embeddings.dtype
        nop, ops_ = cfg.max_output_patches, cfg.output_patch_size
        fut_cov = torch.zeros(b, nop, ops_, dtype=dtype)
        fut_mask = torch.zeros(b, nop, ops_, dtype=dtype)
        fut_time = (
            torch.arange(0, nop * ops_, dtype=torch.float32).div(cfg.time_encoding_scale).reshape(1, nop, ops_)
            .expand(b, -1, -1).to(dtype)
        )
        future_embeds = self._model.input_patch_embedding(torch.cat([fut_time, fut_cov, fut_mask], dim=-1))
        attention_mask = (~masks).to(dtype)
        future_attention_mask = torch.ones(b, nop, dtype=dtype)
        if cfg.use_reg_token:
            reg_ids = torch.full((b, 1), cfg.reg_token_id)
            reg_embeds = self._model.shared(reg_ids)
            input_embeds = torch.cat([input_embeddings, reg_embeds, future_embeds], dim=-2)
            attention_mask = torch.cat([attention_mask, torch.ones_like(reg_ids).to(dtype), future_attention_mask], dim=-1)
        else:
            input_embeds = torch.cat([input_embeddings, future_embeds], dim=-2)
            attention_mask = torch.cat([attention_mask, future_attention_mask], dim=-1)
        group_ids = torch.arange(b, dtype=torch.long)
        hidden = self._model.encoder(inputs_embeds=input_embeds, group_ids=group_ids, attention_mask=attention_mask)[0]
        return hidden[:, -nop:]

    def postprocess(self, horizon, output_embeddings, normalization_stats):  # chronos.py:128-169
        cfg = self._model.cfg
        nop, ops_ = cfg.max_output_patches, cfg.output_patch_size
        max_horizon = nop * ops_
        if horizon > max_horizon:
            raise ValueError(
                f"horizon ({horizon}) exceeds the maximum prediction length "
                f"({max_horizon} = {nop} patches * {ops_} steps)."
            )
        b, nq = output_embeddings.shape[0], self._model.num_quantiles
        loc, scale = normalization_stats["loc"], normalization_stats["scale"]
        preds = self._model.output_patch_embedding(output_embeddings)
        preds = preds.reshape(b, nop, nq, ops_).permute(0, 2, 1, 3).reshape(b, nq, max_horizon)
        preds = self._model.instance_norm_inverse(preds.reshape(b, nq * max_horizon), (loc, scale)).reshape(b, nq, max_horizon)
        return preds[:, :, :horizon].permute(0, 2, 1)

    def freeze_parameters(self) -> None:
        for p in self.parameters():
            p.requires_grad = False

    def unfreeze_parameters(self) -> None:
        for p in self.parameters():
            p.requires_grad = True

    @torch.no_grad()
    def load_upstream_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        """Load a state dict that uses the upstream ``Chronos2Model`` key names (as the product adapter does)."""

        def g(name):
            return sd[name].detach().float().cpu()

        m = self._model
        m.shared.weight.copy_(g("shared.weight"))
        for blk_name, blk in (("input_patch_embedding", m.input_patch_embedding), ("output_patch_embedding", m.output_patch_embedding)):
            for lin in ("hidden_layer", "output_layer", "residual_layer"):
                getattr(blk, lin).weight.copy_(g(f"{blk_name}.{lin}.weight"))
                getattr(blk, lin).bias.copy_(g(f"{blk_name}.{lin}.bias"))
        for i, blk in enumerate(m.blocks):
            pre = f"encoder.block.{i}.layer."
            for j, (attn, ln) in enumerate(((blk.time_attn, blk.time_ln), (blk.group_attn, blk.group_ln))):
                for proj in "qkvo":
                    getattr(attn, proj).weight.copy_(g(f"{pre}{j}.self_attention.{proj}.weight"))
                ln.weight.copy_(g(f"{pre}{j}.layer_norm.weight"))
            blk.wi.weight.copy_(g(pre + "2.mlp.wi.weight"))
            blk.wo.weight.copy_(g(pre + "2.mlp.wo.weight"))
            blk.ff_ln.weight.copy_(g(pre + "2.layer_norm.weight"))
        m.final_layer_norm.weight.copy_(g("encoder.final_layer_norm.weight"))


def oracle_from_product(decoder):
    """Oracle twin (restated Chronos-2 + the reference decoder / fusion control flow) of a product ``MultimodalDecoder``
    built around a ``Chronos2Adapter``, carrying the same weights."""
    from . import timesfm_oracle as O

    module = decoder.adapter._model
    o_adapter = OracleChronos2Adapter(Chronos2Model(Chronos2Config(num_layers=len(module.encoder.block))))
    o_adapter.load_upstream_state_dict({k: v.detach().cpu() for k, v in module.state_dict().items()})
    fus = decoder.fusion
    dims = [l.weight.shape[1] for l in fus.linears()] + [fus.linears()[-1].weight.shape[0]]
    o = O.OracleDecoder(o_adapter, dims[0], len(dims) - 1, dims[1:-1])
    with torch.no_grad():
        for src, dst in zip(fus.linears(), [m for m in o.fusion.projection if isinstance(m, nn.Linear)]):
            dst.weight.copy_(src.weight.detach().float().cpu())
    return o.eval()


def degenerate_group_attention(block: EncoderBlock, ht: torch.Tensor) -> torch.Tensor:
    """With group_ids = arange(B) (chronos.py:117) every series attends only to itself, so the group attention
    sub-layer is exactly ``h + W_o W_v LN(h)`` for every token whose own time-mask is 1."""
    return ht + block.group_attn.o(block.group_attn.v(block.group_ln(ht)))


def synthetic_sqrt_check() -> float:  # tiny helper used by the CPU tests
    return math.sqrt(2.0)

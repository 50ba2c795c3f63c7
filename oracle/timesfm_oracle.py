"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; never imported by the package.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this module, and only as the checker / the CPU baseline.

Restates the reference's TimesFM hot path on the CPU in fp32:

* ``OracleTimesFM2p5Adapter`` follows the control flow of the reference adapter
  (/root/reference/src/tsfmx/tsfm/timesfm.py:47-83 preprocess, :95-98 forward, :116-129 postprocess)
  line by line.  The arithmetic the reference delegates to the un-vendored third-party package
  ``timesfm`` (pinned to git 8a755c9c755fd5b1fe2f0c8af3b86d7a5b846160, /root/reference/uv.lock:1515-1517)
  is taken from its importable twin in `transformers` 5.5.0 (``transformers.models.timesfm2_5``):
  tokenizer ResidualBlock (modeling_timesfm2_5.py:94-115), decoder layer (:351-390), attention
  (:275-348), rotary embedding (:139-234), point head (:668-673).
* ``update_running_stats`` / ``revin`` restate ``timesfm.torch.util`` (HF twin :528-568 / :496-526).
* ``OracleFusion`` / ``OracleDecoder`` restate /root/reference/src/tsfmx/fusion.py:24-47 and
  decoder.py:62-92 so the oracle also runs where /root/reference does not exist (the GPU box).
  In the build container they are checked against the reference's real classes
  (tests/test_oracle_cpu.py).

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path
(/root/reference/tests/ is empty) and ``timesfm`` itself is not installable here, so the oracle is pinned
against (a) the reference's own ``MultimodalDecoder`` / ``MultimodalFusion`` run in this container and
(b) ``TimesFm2_5Model.forward`` of the HF port (max-abs-diff 0 on identical weights).  Upstream
``timesfm`` numerics are therefore "parity unpinned" (see DESIGN.md).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn
from transformers.masking_utils import create_causal_mask
from transformers.models.timesfm2_5 import modeling_timesfm2_5 as hf
from transformers.models.timesfm2_5.configuration_timesfm2_5 import TimesFm2_5Config


@dataclass
class PreprocessResult:
    input_embeddings: torch.Tensor
    masks: torch.Tensor
    normalization_stats: dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# timesfm.torch.util restatements (imported by the reference at timesfm.py:11)
# ------------------------------------------------------------------------------------------------
def update_running_stats(n, mu, sigma, x, mask):
    """Merge the stats of one patch into the running (n, mu, sigma); population variance.
    HF twin: modeling_timesfm2_5.py:528-568."""
    is_valid = (~mask).to(x.dtype)
    inc_n = is_valid.sum(dim=-1)
    inc_n_safe = torch.where(inc_n == 0, torch.ones_like(inc_n), inc_n)
    inc_mu = (x * is_valid).sum(dim=-1) / inc_n_safe
    inc_mu = torch.where(inc_n == 0, torch.zeros_like(inc_mu), inc_mu)
    centered = x - inc_mu.unsqueeze(-1)
    inc_var = ((centered * is_valid) ** 2).sum(dim=-1) / inc_n_safe
    inc_var = torch.where(inc_n == 0, torch.zeros_like(inc_var), inc_var)
    inc_sigma = torch.sqrt(torch.clamp(inc_var, min=0.0))
    new_n = n + inc_n
    new_n_safe = torch.where(new_n == 0, torch.ones_like(new_n), new_n)
    new_mu = (n * mu + inc_mu * inc_n) / new_n_safe
    new_mu = torch.where(new_n == 0, torch.zeros_like(new_mu), new_mu)
    term1 = n * sigma.pow(2)
    term2 = inc_n * inc_sigma.pow(2)
    term3 = n * (mu - new_mu).pow(2)
    term4 = inc_n * (inc_mu - new_mu).pow(2)
    new_var = (term1 + term2 + term3 + term4) / new_n_safe
    new_var = torch.where(new_n == 0, torch.zeros_like(new_var), new_var)
    new_sigma = torch.sqrt(torch.clamp(new_var, min=0.0))
    return (new_n, new_mu, new_sigma), (new_n, new_mu, new_sigma)


def revin(x, mu, sigma, reverse: bool = False):
    """(x - mu) / where(sigma < 1e-6, 1, sigma), or x * sigma + mu.  HF twin: :496-526."""
    if mu.dim() == x.dim() - 1:
        mu, sigma = mu[..., None], sigma[..., None]
    elif mu.dim() == x.dim() - 2:
        mu, sigma = mu[..., None, None], sigma[..., None, None]
    if reverse:
        return x * sigma + mu
    return (x - mu) / torch.where(sigma < 1e-6, torch.ones_like(sigma), sigma)


# ------------------------------------------------------------------------------------------------
# the adapter
# ------------------------------------------------------------------------------------------------
def make_hf_config(num_layers: int) -> TimesFm2_5Config:
    cfg = TimesFm2_5Config(num_hidden_layers=num_layers)
    cfg._attn_implementation = "eager"  # additive finfo.min mask, fp32 softmax
    return cfg


class OracleTimesFM2p5Adapter(nn.Module):
    """Reference ``TimesFM2p5Adapter`` with the third-party module replaced by its HF twin."""

    def __init__(self, num_layers: int = 20, with_quantile_head: bool = False) -> None:
        super().__init__()
        self.cfg = make_hf_config(num_layers)
        self.p, self.o, self.os, self.q, self.md = 32, 128, 1024, 10, 1280
        self.decode_index = 5
        c = self.cfg
        self.tokenizer = hf.TimesFm2_5ResidualBlock(c, 2 * self.p, self.md, self.md, use_bias=True)
        self.stacked_xf = nn.ModuleList(hf.TimesFm2_5DecoderLayer(c, i) for i in range(num_layers))
        self.rotary_emb = hf.TimesFm2_5RotaryEmbedding(c)
        self.output_projection_point = hf.TimesFm2_5ResidualBlock(c, self.md, self.md, self.o * self.q)
        if with_quantile_head:  # continuous quantile head (HF :676-681); unused by the reference adapter
            self.output_projection_quantiles = hf.TimesFm2_5ResidualBlock(c, self.md, self.md, self.os * self.q)
        self.eval()

    @property
    def model_dims(self) -> int:
        return self.md

    @property
    def patch_len(self) -> int:
        return self.p

    @property
    def point_forecast_index(self) -> int:
        return self.decode_index

    # reference timesfm.py:36-83
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult:
        batch_size, context = inputs.shape[0], inputs.shape[1]
        if context % self.p != 0:
            raise ValueError(f"context length ({context}) must be divisible by patch length ({self.p})")
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        num_input_patches = context // self.p
        patched_inputs = inputs.reshape(batch_size, -1, self.p)
        patched_masks = masks.reshape(batch_size, -1, self.p)
        n = torch.zeros(batch_size)
        mu = torch.zeros(batch_size)
        sigma = torch.zeros(batch_size)
        patch_mu, patch_sigma = [], []
        for i in range(num_input_patches):
            (n, mu, sigma), _ = update_running_stats(n, mu, sigma, patched_inputs[:, i], patched_masks[:, i])
            patch_mu.append(mu)
            patch_sigma.append(sigma)
        context_mu = torch.stack(patch_mu, dim=1)
        context_sigma = torch.stack(patch_sigma, dim=1)
        normed_inputs = revin(patched_inputs, context_mu, context_sigma, reverse=False)
        normed_inputs = torch.where(patched_masks, 0.0, normed_inputs)
        tokenizer_inputs = torch.cat([normed_inputs, patched_masks.to(normed_inputs.dtype)], dim=-1)
        input_embeddings = self.tokenizer(tokenizer_inputs)
        return PreprocessResult(
            input_embeddings=input_embeddings,
            masks=patched_masks,
            normalization_stats={"context_mu": context_mu, "context_sigma": context_sigma},
        )

    # reference timesfm.py:85-98; `layer(x, masks[..., -1], None)` of upstream == HF layer driven as in
    # TimesFm2_5Model.forward (:620-641): position = arange(N) - num_masked, causal & key-not-padded mask.
    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        patch_padding = masks[..., -1]
        seq = input_embeddings.shape[1]
        num_masked = patch_padding.to(torch.int32).sum(dim=-1, keepdim=True)
        position_ids = torch.arange(seq).unsqueeze(0) - num_masked
        padding_mask = (~patch_padding).to(torch.int64)
        attention_mask = create_causal_mask(self.cfg, input_embeddings, padding_mask, past_key_values=None)
        position_embeddings = self.rotary_emb(input_embeddings, position_ids)
        output_embeddings = input_embeddings
        for layer in self.stacked_xf:
            output_embeddings = layer(
                output_embeddings,
                position_embeddings=position_embeddings,
                attention_mask=attention_mask,
                position_ids=position_ids,
            )
        return output_embeddings

    # reference timesfm.py:100-129
    def postprocess(self, horizon: int, output_embeddings: torch.Tensor, normalization_stats) -> torch.Tensor:
        if horizon > self.o:
            raise ValueError(
                f"horizon must be <= output_patch_len ({self.o}), got {horizon}. AR decode is not supported."
            )
        batch_size = output_embeddings.shape[0]
        context_mu = normalization_stats["context_mu"]
        context_sigma = normalization_stats["context_sigma"]
        output_ts = self.output_projection_point(output_embeddings)
        renormed = torch.reshape(
            revin(output_ts, context_mu, context_sigma, reverse=True), (batch_size, -1, self.o, self.q)
        )
        return renormed[:, -1, :horizon, :]

    def freeze_parameters(self) -> None:
        for p in self.parameters():
            p.requires_grad = False

    def unfreeze_parameters(self) -> None:
        for p in self.parameters():
            p.requires_grad = True

    # ------------------------------------------------------------------------------------------------
    # beyond the reference adapter: upstream's decode loop for horizons > 128 and its
    # forecast extras.  PARITY STATUS: the extras (continuous quantile head, flip invariance, positivity clamp) are
    # pinned to HF ``TimesFm2_5ModelForPrediction.forward`` (modeling_timesfm2_5.py:797-837, tests/test_oracle_cpu.py);
    # the autoregressive loop is RESTATED from upstream ``TimesFM_2p5_200M_torch_module.decode`` (timesfm @ 8a755c9,
    # not in this container; neither the reference nor the HF port implements it) -> "parity unpinned" for h > 128.
    # ------------------------------------------------------------------------------------------------
    def _running_stats(self, patched_inputs, patched_masks, state=None):
        """Per-patch running (mu, sigma) continuing from ``state`` = (n, mu, sigma); -> (mu [B,N], sigma [B,N], state)."""
        b = patched_inputs.shape[0]
        n, mu, sigma = state if state is not None else (torch.zeros(b), torch.zeros(b), torch.zeros(b))
        mus, sigmas = [], []
        for i in range(patched_inputs.shape[1]):
            (n, mu, sigma), _ = update_running_stats(n, mu, sigma, patched_inputs[:, i], patched_masks[:, i])
            mus.append(mu)
            sigmas.append(sigma)
        return torch.stack(mus, dim=1), torch.stack(sigmas, dim=1), (n, mu, sigma)

    def _tokenize(self, patched_inputs, patched_masks, mu, sigma):
        normed = revin(patched_inputs, mu, sigma, reverse=False)
        normed = torch.where(patched_masks, 0.0, normed)
        return self.tokenizer(torch.cat([normed, patched_masks.to(normed.dtype)], dim=-1))

    def _stack(self, embeddings: torch.Tensor, patch_padding: torch.Tensor) -> torch.Tensor:
        return self.forward(embeddings, patch_padding[..., None])  # forward() reads masks[..., -1]

    def decode(self, horizon: int, inputs: torch.Tensor, masks: torch.Tensor, fuse=None):
        """Upstream ``decode``: prefill on the context, then ``(horizon - 1) // 128`` autoregressive steps, each feeding
        the previous 128-step point forecast (channel ``decode_index``) back as 4 new patches whose running statistics
        continue the context's.  A causal stack gives an appended token the same output whether the prefix is cached
        or recomputed, so the oracle simply re-runs the whole sequence every step (the product keeps a KV cache).

        ``fuse(embeddings) -> embeddings`` is applied to the CONTEXT patches only (text exists for them alone).
        Returns (point/quantile forecast [B, 128 * (steps + 1), 10], quantile spread [B, 1024, 10] or None)."""
        b, context = inputs.shape
        if context % self.p != 0:
            raise ValueError(f"context length ({context}) must be divisible by patch length ({self.p})")
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        patched_inputs, patched_masks = inputs.reshape(b, -1, self.p), masks.reshape(b, -1, self.p)
        mu, sigma, state = self._running_stats(patched_inputs, patched_masks)
        emb = self._tokenize(patched_inputs, patched_masks, mu, sigma)
        if fuse is not None:
            emb = fuse(emb)
        padding = patched_masks[..., -1]
        out = self._stack(emb, padding)
        pf = revin(self.output_projection_point(out[:, -1]), mu[:, -1:], sigma[:, -1:], reverse=True).reshape(b, self.o, self.q)
        spread = None
        if hasattr(self, "output_projection_quantiles"):
            spread = revin(self.output_projection_quantiles(out[:, -1]), mu[:, -1:], sigma[:, -1:], reverse=True)
            spread = spread.reshape(b, self.os, self.q)
        outputs = [pf]
        m = self.o // self.p
        for _ in range((horizon - 1) // self.o):
            new_inputs = outputs[-1][:, :, self.decode_index].reshape(b, m, self.p)
            new_masks = torch.zeros_like(new_inputs, dtype=torch.bool)
            new_mu, new_sigma, state = self._running_stats(new_inputs, new_masks, state)
            emb = torch.cat([emb, self._tokenize(new_inputs, new_masks, new_mu, new_sigma)], dim=1)
            padding = torch.cat([padding, torch.zeros(b, m, dtype=torch.bool)], dim=1)
            out = self._stack(emb, padding)
            step = revin(self.output_projection_point(out[:, -1]), new_mu[:, -1:], new_sigma[:, -1:], reverse=True)
            outputs.append(step.reshape(b, self.o, self.q))
        return torch.cat(outputs, dim=1), spread

    # ---- weights: from the product's (upstream-named) state dict
    @torch.no_grad()
    def load_upstream_state_dict(self, sd: dict[str, torch.Tensor], prefix: str = "") -> None:
        def g(name: str) -> torch.Tensor:
            return sd[prefix + name].detach().float().cpu()

        for src, dst in (("hidden_layer", "input_layer"), ("output_layer", "output_layer"), ("residual_layer", "residual_layer")):
            getattr(self.tokenizer, dst).weight.copy_(g(f"tokenizer.{src}.weight"))
            getattr(self.tokenizer, dst).bias.copy_(g(f"tokenizer.{src}.bias"))
            getattr(self.output_projection_point, dst).weight.copy_(g(f"output_projection_point.{src}.weight"))
            if hasattr(self, "output_projection_quantiles"):
                getattr(self.output_projection_quantiles, dst).weight.copy_(g(f"output_projection_quantiles.{src}.weight"))
        d = self.md
        for i, layer in enumerate(self.stacked_xf):
            pre = f"stacked_xf.{i}."
            layer.input_layernorm.weight.copy_(g(pre + "pre_attn_ln.scale"))
            layer.post_attention_layernorm.weight.copy_(g(pre + "post_attn_ln.scale"))
            layer.pre_feedforward_layernorm.weight.copy_(g(pre + "pre_ff_ln.scale"))
            layer.post_feedforward_layernorm.weight.copy_(g(pre + "post_ff_ln.scale"))
            qkv = g(pre + "attn.qkv_proj.weight")
            layer.self_attn.q_proj.weight.copy_(qkv[:d])
            layer.self_attn.k_proj.weight.copy_(qkv[d : 2 * d])
            layer.self_attn.v_proj.weight.copy_(qkv[2 * d :])
            layer.self_attn.o_proj.weight.copy_(g(pre + "attn.out.weight"))
            layer.self_attn.q_norm.weight.copy_(g(pre + "attn.query_ln.scale"))
            layer.self_attn.k_norm.weight.copy_(g(pre + "attn.key_ln.scale"))
            layer.self_attn.scaling.copy_(g(pre + "attn.per_dim_scale.per_dim_scale"))
            layer.mlp.fc1.weight.copy_(g(pre + "ff0.weight"))
            layer.mlp.fc2.weight.copy_(g(pre + "ff1.weight"))


# ------------------------------------------------------------------------------------------------
# fusion + decoder restatements (reference fusion.py:24-47, decoder.py:62-92)
# ------------------------------------------------------------------------------------------------
class OracleFusion(nn.Module):
    def __init__(self, ts_dims: int, text_dims: int, num_layers: int = 1, hidden_dims: list[int] | None = None):
        super().__init__()
        hidden_dims = list(hidden_dims or [])
        if num_layers < 1 or num_layers > 3:
            raise ValueError(f"num_layers must be between 1 and 3, got {num_layers}")
        if len(hidden_dims) != num_layers - 1:
            raise ValueError(
                f"hidden_dims must have {num_layers - 1} elements for {num_layers} layers, got {len(hidden_dims)}"
            )
        dims = [text_dims, *hidden_dims, ts_dims]
        layers: list[nn.Module] = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1], bias=False))
            layers.append(nn.ReLU())  # after EVERY Linear, including the last (fusion.py:27-29)
        self.projection = nn.Sequential(*layers)

    def forward(self, ts_embeddings, text_embeddings):
        return ts_embeddings + self.projection(text_embeddings)


@dataclass
class ForecastOptions:
    """Upstream ``ForecastConfig`` switches (HF config names :87-89).  All off = the reference adapter's behaviour."""

    use_continuous_quantile_head: bool = False
    force_flip_invariance: bool = False
    infer_is_positive: bool = False


def flip_quantiles(x: torch.Tensor) -> torch.Tensor:
    """Channel 0 (mean) stays, the nine quantile channels are reversed (HF :800-801)."""
    return torch.cat([x[..., :1], torch.flip(x[..., 1:], dims=(-1,))], dim=-1)


def apply_forecast_extras(pf, spread, pf_neg, spread_neg, inputs, horizon, options: ForecastOptions, decode_index=5):
    """HF ``TimesFm2_5ModelForPrediction.forward`` :797-837 on already decoded outputs: flip-invariance combination,
    continuous quantile head, horizon slice, positivity clamp (per series, as upstream; HF clamps on the batch minimum,
    which coincides whenever all series of the batch agree)."""
    if options.force_flip_invariance:
        pf = (pf - flip_quantiles(pf_neg)) / 2
        if spread is not None:
            spread = (spread - flip_quantiles(spread_neg)) / 2
    full = pf[:, :horizon, :].clone()
    if options.use_continuous_quantile_head:
        hq = min(horizon, spread.shape[1])
        for idx in range(1, full.shape[-1]):
            if idx == decode_index:
                continue
            full[:, :hq, idx] = spread[:, :hq, idx] - spread[:, :hq, decode_index] + full[:, :hq, decode_index]
    if options.infer_is_positive:
        positive = (inputs.min(dim=-1).values >= 0)[:, None, None]
        full = torch.where(positive, full.clamp_min(0.0), full)
    return full


class OracleDecoder(nn.Module):
    def __init__(self, adapter: nn.Module, text_dims: int = 384, num_layers: int = 1, hidden_dims=None):
        super().__init__()
        self.adapter = adapter
        self.fusion = OracleFusion(adapter.model_dims, text_dims, num_layers, hidden_dims)

    def forward_full(self, horizon, inputs, masks, text_embeddings=None):
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        masks = masks.bool()
        pre = self.adapter.preprocess(inputs, masks)
        emb = self.fusion(pre.input_embeddings, text_embeddings) if text_embeddings is not None else pre.input_embeddings
        out = self.adapter(emb, pre.masks)
        return self.adapter.postprocess(horizon, out, pre.normalization_stats)

    def forward(self, horizon, inputs, masks, text_embeddings=None):
        return self.forward_full(horizon, inputs, masks, text_embeddings)[..., self.adapter.point_forecast_index]

    def forecast(self, horizon, inputs, masks, text_embeddings=None, options: ForecastOptions | None = None):
        """``forward_full`` through upstream's decode loop: any horizon (AR steps of 128) + the forecast extras.  With
        horizon <= 128 and default options it returns exactly ``forward_full``."""
        options = options or ForecastOptions()
        masks = masks.bool()
        fuse = (lambda e: self.fusion(e, text_embeddings)) if text_embeddings is not None else None
        pf, spread = self.adapter.decode(horizon, inputs, masks, fuse)
        pf_neg = spread_neg = None
        if options.force_flip_invariance:
            pf_neg, spread_neg = self.adapter.decode(horizon, -inputs, masks, fuse)
        return apply_forecast_extras(pf, spread, pf_neg, spread_neg, inputs, horizon, options, self.adapter.decode_index)


def oracle_from_product(decoder) -> OracleDecoder:
    """Build an oracle decoder carrying the exact weights of a product ``MultimodalDecoder`` (TimesFM adapter)."""
    adapter = decoder.adapter
    num_layers = len(adapter._model.stacked_xf)
    o_adapter = OracleTimesFM2p5Adapter(num_layers, with_quantile_head=hasattr(adapter._model, "output_projection_quantiles"))
    o_adapter.load_upstream_state_dict(adapter._model.state_dict())
    fus = decoder.fusion
    dims = [l.weight.shape[1] for l in fus.linears()] + [fus.linears()[-1].weight.shape[0]]
    o = OracleDecoder(o_adapter, dims[0], len(dims) - 1, dims[1:-1])
    with torch.no_grad():
        for src, dst in zip(fus.linears(), [m for m in o.fusion.projection if isinstance(m, nn.Linear)]):
            dst.weight.copy_(src.weight.detach().float().cpu())
    return o.eval()


# ------------------------------------------------------------------------------------------------
# synthetic Time-MMD-shaped inputs
# ------------------------------------------------------------------------------------------------
def synthetic_batch(batch: int, context: int, horizon: int, text_dims: int = 384, patch_len: int = 32, seed: int = 1234,
                    padded: bool = False):
    """context (B,C) z-normalised per row, masks (B,C) bool, text (B,N,E) unit-norm rows, horizon (B,h)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(context + horizon, dtype=torch.float32)
    raw = torch.randn(batch, context + horizon, generator=g).cumsum(-1) * 0.3
    raw = raw + 2.0 * torch.sin(2 * math.pi * t / 24.0 + torch.rand(batch, 1, generator=g) * 6.28)
    ctx, hor = raw[:, :context], raw[:, context:]
    mean = ctx.mean(-1, keepdim=True)
    std = ctx.std(-1, keepdim=True)
    std = torch.where(std < 1e-6, torch.ones_like(std), std)
    ctx, hor = (ctx - mean) / std, (hor - mean) / std
    masks = torch.zeros(batch, context, dtype=torch.bool)
    if padded:
        gp = torch.Generator().manual_seed(seed + 1)
        pad = torch.randint(0, context - patch_len, (batch,), generator=gp)
        pad[0] = 0
        if batch > 1:
            pad[1] = patch_len + 5  # one full patch + a partial one
        if batch > 2:
            pad[2] = 3 * patch_len
        masks = torch.arange(context)[None, :] < pad[:, None]
    gt = torch.Generator().manual_seed(4321 + seed)
    text = torch.randn(batch, context // patch_len, text_dims, generator=gt)
    text = text / text.norm(dim=-1, keepdim=True)
    return ctx.contiguous(), masks, text.contiguous(), hor.contiguous()


# ------------------------------------------------------------------------------------------------
# the oracle itself in bf16 (calibrates the tolerance of the product's "bf16" throughput mode)
# ------------------------------------------------------------------------------------------------
_BF16_OUTPUT_LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj", "fc2")


def bf16_oracle(oracle: nn.Module, output_linears: tuple[str, ...] = _BF16_OUTPUT_LINEARS) -> nn.Module:
    """A copy of ``oracle`` that computes the way a bf16 deployment of the reference would: every ``nn.Linear`` sees
    bf16-rounded weights and bf16-rounded inputs and accumulates in fp32; the outputs of the attention projections and
    of the second MLP matrix are rounded to bf16 as well (they are stored as bf16 activations).  Everything else -
    statistics, norms, softmax, residual stream - stays fp32, exactly as in the fp32 oracle.

    Its distance to the fp32 oracle is what bf16 operands cost on THIS model and THESE inputs, independent of any
    kernel: SURVEY.md section 8(d) asks for the bf16 tolerance to be "calibrated against the oracle itself run with
    bf16 weights/activations".  tests/test_parity_gpu.py and bench.py bound the product's bf16 error by a stated
    multiple of it instead of a hand-set constant.

    ``output_linears``: last name components (or dotted suffixes) of the Linears whose OUTPUT is stored in bf16 by a
    bf16 deployment - TimesFM's by default; ``CHRONOS2_BF16_OUTPUTS`` / ``T5_BF16_OUTPUTS`` for the other oracles."""
    import copy

    twin = copy.deepcopy(oracle).eval()

    def to_bf16(t: torch.Tensor) -> torch.Tensor:
        return t.to(torch.bfloat16).to(torch.float32)

    def round_inputs(_module, args):
        return tuple(to_bf16(a) if isinstance(a, torch.Tensor) and a.is_floating_point() else a for a in args)

    def round_output(_module, _args, out):
        return to_bf16(out)

    with torch.no_grad():
        for name, module in twin.named_modules():
            if isinstance(module, nn.Linear):
                module.weight.copy_(to_bf16(module.weight))
                module.register_forward_pre_hook(round_inputs)
                if name.rsplit(".", 1)[-1] in output_linears or name.endswith(tuple("." + o for o in output_linears if "." in o)):
                    module.register_forward_hook(round_output)
    return twin


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a - b| / max|b|: the relative error the north star's 1e-3 bar is stated in."""
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


CHRONOS2_BF16_OUTPUTS = ("time_attn.q", "time_attn.k", "time_attn.v", "time_attn.o", "group_attn.o", "wo")
T5_BF16_OUTPUTS = ("q", "k", "v", "o", "wo")

# the product's bf16 mode may deviate from the fp32 oracle by at most this multiple of the bf16 oracle's own deviation
# (two independent bf16 realisations of the same computation differ from each other by about sqrt(2) of it; the product
# additionally keeps P, q' and k' of the attention core in bf16 for the tensor cores)
BF16_TOL_FACTOR = 3.0

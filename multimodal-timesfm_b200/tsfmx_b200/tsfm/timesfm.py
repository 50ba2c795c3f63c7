"""TimesFM 2.5 adapter (reference tsfmx/tsfm/timesfm.py:17-166), B200-native.

The reference wraps the third-party ``timesfm`` torch module; here the parameters live in a plain
container with the upstream state-dict key names (so ``google/timesfm-2.5-200m-pytorch`` safetensors
load with ``strict=True``) and every stage runs through the C ABI:

  preprocess  : fused patchify + running RevIN stats + mask-concat kernel, then the tokenizer
                ResidualBlock as two tcgen05 GEMMs (bias + SiLU epilogue; hidden and residual paths
                accumulated into one TMEM tile)
  forward     : per layer  QKV GEMM -> attention kernel -> out-proj GEMM -> fused post-norm +
                residual + pre-norm -> ff0 GEMM (SiLU epilogue) -> ff1 GEMM -> fused norms
  postprocess : point head on the LAST patch only (the reference projects all N patches and keeps
                ``[:, -1]``, timesfm.py:125-129) with inverse RevIN + horizon slice in the epilogue
"""

from __future__ import annotations

import os
import math
from dataclasses import dataclass
from types import SimpleNamespace

import torch
from torch import nn

from .. import ops
from .._lib import ACT_SILU, ACT_SILU_GRAD, DT_BF16, DT_F32, MAX_KV_REGIONS, PREC_BF16, PRECISIONS, TsfmxError
from ..lanes import drain
from .base import PreprocessResult, TsfmAdapter


# --------------------------------------------------------------------------- parameter containers
class _ResidualBlock(nn.Module):
    """output_layer(act(hidden_layer(x))) + residual_layer(x); parameters only."""

    def __init__(self, d_in: int, d_hidden: int, d_out: int, bias: bool) -> None:
        super().__init__()
        self.hidden_layer = nn.Linear(d_in, d_hidden, bias=bias)
        self.output_layer = nn.Linear(d_hidden, d_out, bias=bias)
        self.residual_layer = nn.Linear(d_in, d_out, bias=bias)


class _RMSNorm(nn.Module):
    def __init__(self, dims: int) -> None:
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dims))


class _PerDimScale(nn.Module):
    def __init__(self, dims: int) -> None:
        super().__init__()
        self.per_dim_scale = nn.Parameter(torch.zeros(dims))


class _Attention(nn.Module):
    def __init__(self, md: int, heads: int, hd: int) -> None:
        super().__init__()
        self.qkv_proj = nn.Linear(md, 3 * heads * hd, bias=False)  # rows: [q | k | v], head-major inside
        self.out = nn.Linear(heads * hd, md, bias=False)
        self.query_ln = _RMSNorm(hd)
        self.key_ln = _RMSNorm(hd)
        self.per_dim_scale = _PerDimScale(hd)


class _Transformer(nn.Module):
    def __init__(self, md: int, heads: int, hd: int, ff: int) -> None:
        super().__init__()
        self.pre_attn_ln = _RMSNorm(md)
        self.post_attn_ln = _RMSNorm(md)
        self.attn = _Attention(md, heads, hd)
        self.pre_ff_ln = _RMSNorm(md)
        self.post_ff_ln = _RMSNorm(md)
        self.ff0 = nn.Linear(md, ff, bias=False)
        self.ff1 = nn.Linear(ff, md, bias=False)


class TimesFM2p5Module(nn.Module):
    """Parameters of TimesFM 2.5 with upstream attribute / key names (p, o, os, q, md, x, ...)."""

    def __init__(self, num_layers: int = 20, with_quantile_head: bool = True) -> None:
        super().__init__()
        self.p, self.o, self.os, self.q = 32, 128, 1024, 10
        self.md, self.h, self.hd, self.ff, self.x = 1280, 16, 80, 1280, num_layers
        self.eps = 1e-6
        self.config = SimpleNamespace(decode_index=5, context_limit=16384)
        self.tokenizer = _ResidualBlock(2 * self.p, self.md, self.md, bias=True)
        self.stacked_xf = nn.ModuleList(_Transformer(self.md, self.h, self.hd, self.ff) for _ in range(num_layers))
        self.output_projection_point = _ResidualBlock(self.md, self.md, self.o * self.q, bias=False)
        if with_quantile_head:  # unused by tsfmx; kept so upstream checkpoints load strictly
            self.output_projection_quantiles = _ResidualBlock(self.md, self.md, self.os * self.q, bias=False)


def init_random_(model: nn.Module, seed: int = 0) -> None:
    """Deterministic random init of EVERY tensor: Linear ~ N(0, 0.02), biases ~ N(0, 0.01),
    RMSNorm scales 1 + 0.1 N(0, 1), per-dim scale ~ N(0, 0.5).  Generated on the CPU so the values do not
    depend on the device."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("per_dim_scale"):
                v = 0.5 * torch.randn(p.shape, generator=gen)
            elif name.endswith(".scale"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=gen)
            elif name.endswith(".bias"):
                v = 0.01 * torch.randn(p.shape, generator=gen)
            else:
                v = 0.02 * torch.randn(p.shape, generator=gen)
            p.copy_(v.to(p.device))


@dataclass
class ForecastOptions:
    """What upstream timesfm does around the stack and the reference adapter leaves out.
    Everything off (the default) is exactly the reference: point head only, ``ValueError`` for horizon > 128.

    ``ar_decode``: horizons above ``output_patch_len`` are decoded autoregressively, 128 steps at a time, against a KV
    cache (upstream ``decode``).  The other three are upstream's ``ForecastConfig`` switches as the HF port spells them
    (transformers configuration_timesfm2_5.py:87-89)."""

    ar_decode: bool = False
    use_continuous_quantile_head: bool = False
    force_flip_invariance: bool = False
    infer_is_positive: bool = False

    def active(self, horizon: int, output_patch_len: int) -> bool:
        return (self.ar_decode and horizon > output_patch_len) or self.use_continuous_quantile_head \
            or self.force_flip_invariance or self.infer_is_positive


# --------------------------------------------------------------------------- adapter
class TimesFM2p5Adapter(TsfmAdapter):
    """Adapter for TimesFM 2.5 (200 M layout; ``num_layers=50`` gives the "500 M" shape of BASELINE.json)."""

    def __init__(self, num_layers: int = 20, precision: str = "bf16", with_quantile_head: bool = True) -> None:
        super().__init__()
        self._model = TimesFM2p5Module(num_layers, with_quantile_head)
        self.fused_norm = os.environ.get("TSFMX_FUSED_NORM", "0") == "1"  # True: norm/residual junctions in the GEMM epilogue (5-CTA clusters; measured slower, kept for A/B)
        self.set_precision(precision)
        self._packed: dict[object, dict[str, object]] = {}
        self.forecast_options = ForecastOptions()
        # forecast through the whole-stack entry point (one library call for all layers); False = one call per kernel,
        # which lets two series lanes interleave their launches kernel by kernel
        self.stack_call = os.environ.get("TSFMX_STACK_CALL", "1") == "1"
        # RoPE frequencies: computed once on the CPU (bit-identical to the reference's), moved with the module - a
        # host-to-device copy inside _weights() would be illegal while a training step is being captured into a graph
        hd = self._model.hd
        self.register_buffer("_inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)),
                             persistent=False)

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision

    @property
    def model_dims(self) -> int:
        return int(self._model.md)

    @property
    def patch_len(self) -> int:
        return int(self._model.p)

    @property
    def num_outputs(self) -> int:
        """Channels of the ``postprocess`` output (not part of the reference contract; used for empty batches)."""
        return int(self._model.q)

    @property
    def point_forecast_index(self) -> int:
        return int(self._model.config.decode_index)

    # ------------------------------------------------------------------ packed weights
    def _weights(self) -> dict[str, object]:
        """bf16 (or split-bf16) copies of the Linear weights + derived per-layer vectors, cached until a
        parameter changes."""
        prec = PRECISIONS[self.precision]
        params = list(self._model.parameters())
        key = (prec, tuple((p.data_ptr(), p._version) for p in params))
        if key in self._packed:
            return self._packed[key]
        self._packed.clear()
        m = self._model
        adt = ops.act_dtype(prec)

        def pack(lin: nn.Linear) -> torch.Tensor:
            return ops.cast_rows(lin.weight.detach().float().contiguous(), adt)

        def pack_t(lin: nn.Linear) -> torch.Tensor:
            """W^T, K-major: the B operand of the dgrad GEMM dX = dY W (built lazily, only for training)."""
            wt = lin.weight.detach()
            n_out, k_in = wt.shape
            if wt.dtype == torch.float32 and wt.is_contiguous() and n_out % 64 == 0:
                return ops.transpose_mask(wt, n_out, k_in, adt)[0]  # transpose + cast in one kernel
            return ops.cast_rows(wt.float().t().contiguous(), adt)

        def f32(t: torch.Tensor) -> torch.Tensor:
            return t.detach().float().contiguous()

        dev = params[0].device
        hd = m.hd
        w: dict[str, object] = {
            "tok_hidden": pack(m.tokenizer.hidden_layer),
            "tok_hidden_b": f32(m.tokenizer.hidden_layer.bias),
            "tok_out": pack(m.tokenizer.output_layer),
            "tok_res": pack(m.tokenizer.residual_layer),
            "tok_out_b": f32(m.tokenizer.output_layer.bias + m.tokenizer.residual_layer.bias),
            "head_hidden": pack(m.output_projection_point.hidden_layer),
            "head_out": pack(m.output_projection_point.output_layer),
            "head_res": pack(m.output_projection_point.residual_layer),
            "inv_freq": self._inv_freq.to(dev),
            "layers": [],
        }
        if hasattr(m, "output_projection_quantiles"):  # continuous quantile head (ForecastOptions)
            w["qhead_hidden"] = pack(m.output_projection_quantiles.hidden_layer)
            w["qhead_out"] = pack(m.output_projection_quantiles.output_layer)
            w["qhead_res"] = pack(m.output_projection_quantiles.residual_layer)
        for xf in m.stacked_xf:
            w["layers"].append(
                {
                    "pre_attn": f32(xf.pre_attn_ln.scale),
                    "post_attn": f32(xf.post_attn_ln.scale),
                    "pre_ff": f32(xf.pre_ff_ln.scale),
                    "post_ff": f32(xf.post_ff_ln.scale),
                    "qkv": pack(xf.attn.qkv_proj),
                    "out": pack(xf.attn.out),
                    "q_ln": f32(xf.attn.query_ln.scale),
                    "k_ln": f32(xf.attn.key_ln.scale),
                    "q_scale": f32(
                        torch.nn.functional.softplus(xf.attn.per_dim_scale.per_dim_scale.detach().float())
                        * (1.442695041 / math.sqrt(hd))
                    ),
                    "ff0": pack(xf.ff0),
                    "ff1": pack(xf.ff1),
                }
            )
        w["_pack_t"] = pack_t
        self._packed[key] = w
        return w

    def _weights_t(self) -> dict[str, object]:
        """Transposed (dgrad) packs of the frozen backbone weights, cached next to the forward packs."""
        w = self._weights()
        if "t" not in w:
            m, pack_t = self._model, w["_pack_t"]
            w["t"] = {
                "tok_out": pack_t(m.tokenizer.output_layer),
                "head_hidden": pack_t(m.output_projection_point.hidden_layer),
                "head_out": pack_t(m.output_projection_point.output_layer),
                "head_res": pack_t(m.output_projection_point.residual_layer),
                "layers": [
                    {"qkv": pack_t(xf.attn.qkv_proj), "out": pack_t(xf.attn.out), "ff0": pack_t(xf.ff0),
                     "ff1": pack_t(xf.ff1)}
                    for xf in m.stacked_xf
                ],
            }
        return w["t"]

    # ------------------------------------------------------------------ stages
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult:
        """Patch, normalise (running RevIN) and tokenize (reference timesfm.py:36-83).

        Raises ValueError if the context is not a multiple of the patch length or the mask shape differs.
        """
        return drain(self.preprocess_steps(inputs, masks))

    def preprocess_steps(self, inputs: torch.Tensor, masks: torch.Tensor):
        """``preprocess`` as a step generator (one ``yield`` per kernel launch, see ``tsfmx_b200.lanes``)."""
        m = self._model
        batch_size, context = inputs.shape[0], inputs.shape[1]
        if context % m.p != 0:
            raise ValueError(f"context length ({context}) must be divisible by patch length ({m.p})")
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        if not inputs.is_cuda:
            raise TsfmxError("TimesFM2p5Adapter runs on B200 only; there is no CPU fallback")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        n = context // m.p
        rows = batch_size * n
        masks = masks.bool()
        tokens, mu, sigma, _pm, _nm = ops.timesfm_patchify_norm(inputs, masks, m.p, adt)
        yield
        hidden = ops.alloc(rows, m.md, adt, inputs.device)
        ops.gemm([(tokens, w["tok_hidden"], 2 * m.p)], rows, m.md, hidden, adt, precision=prec, act=ACT_SILU,
                 bias=w["tok_hidden_b"])
        yield
        emb = torch.empty(rows, m.md, dtype=torch.float32, device=inputs.device)
        ops.gemm([(hidden, w["tok_out"], m.md), (tokens, w["tok_res"], 2 * m.p)], rows, m.md, emb, DT_F32,
                 precision=prec, bias=w["tok_out_b"])
        yield
        return PreprocessResult(
            input_embeddings=emb.view(batch_size, n, m.md),
            masks=masks.reshape(batch_size, n, m.p),
            normalization_stats={"context_mu": mu, "context_sigma": sigma},
        )

    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """Run the stacked transformer layers (reference timesfm.py:85-98) -> (batch, patches, model_dims)."""
        return drain(self.forward_steps(input_embeddings, masks))

    def forward_steps(self, input_embeddings: torch.Tensor, masks: torch.Tensor, kv_cache: list | None = None):
        """``forward`` as a step generator (one ``yield`` per kernel launch, see ``tsfmx_b200.lanes``).

        ``kv_cache``: an empty list to be filled with one ``[qkv]`` region list per layer - the raw qkv matrices are
        then kept (one buffer per layer instead of one reused scratch) for the decode steps that follow."""
        if not input_embeddings.is_cuda:
            raise TsfmxError("TimesFM2p5Adapter runs on B200 only; there is no CPU fallback")
        b, n, d = input_embeddings.shape
        x = input_embeddings.reshape(b * n, d).float().contiguous()
        patch_mask = masks[..., -1].contiguous()  # a patch is padded iff its last step is (timesfm.py:97)
        num_masked = patch_mask.sum(-1, dtype=torch.int32)
        y = yield from self._stack_steps(x, b, n, patch_mask, num_masked, kv_cache, decode=False)
        return y.view(b, n, d)

    def _stack_steps(self, x: torch.Tensor, b: int, n: int, patch_mask, num_masked, kv_cache: list | None, decode: bool):
        """The decoder layers over ``n`` tokens per series, x [b * n, D] fp32 -> [b * n, D] fp32.  Prefill
        (``decode=False``): causal attention among the n tokens.  Decode step: the n (= 4) new tokens attend to every
        token already in ``kv_cache`` and to each other."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        fast = prec == PREC_BF16
        mid_dt = DT_BF16 if fast else DT_F32  # GEMM outputs consumed by the norm / attention kernels
        w = self._weights()
        rows, d = x.shape
        dev = x.device
        layers = w["layers"]
        if not layers:
            return x.clone()
        if self.stack_call and not decode and kv_cache is None and not self.fused_norm:
            # one FFI crossing for all layers (tsfmx_timesfm_stack_fwd): same kernels, same order, ~36 us of Python per
            # launch saved (13 ms per 50-layer forward when nothing replays a graph)
            if "stack_table" not in w:
                w["stack_table"] = ops.timesfm_stack_table(layers, w["inv_freq"], m.md, m.h, m.hd, m.ff, prec, m.eps)
            y = ops.timesfm_stack_fwd(w["stack_table"], x, b, n, patch_mask, num_masked)
            yield
            return y
        y = torch.empty(rows, d, dtype=torch.float32, device=dev)
        if kv_cache is not None and not kv_cache:
            kv_cache.extend([] for _ in layers)
        xn = ops.rmsnorm(x, layers[0]["pre_attn"], m.eps, adt)
        yield
        # one scratch qkv when nothing is kept; with a KV cache one allocation holds every layer's qkv (50 x 1 GB at
        # 2048 series of 64 patches: a single block that the caching allocator hands back whole on the next forecast,
        # instead of 50 blocks that fragment against the activations of the other lane)
        qkv = ops.alloc(rows, 3 * d, mid_dt, dev) if kv_cache is None else None
        qkv_all = None if kv_cache is None else ops.alloc(len(layers) * rows, 3 * d, mid_dt, dev)
        attn = ops.alloc(rows, d, adt, dev)
        a = ops.alloc(rows, d, mid_dt, dev)
        hbuf = ops.alloc(rows, m.ff, adt, dev)
        for i, lw in enumerate(layers):
            last = i == len(layers) - 1
            nxt = None if last else layers[i + 1]["pre_attn"]
            if kv_cache is not None:
                qkv = qkv_all[i * rows : (i + 1) * rows]
                kv_cache[i].append(qkv)
            ops.gemm([(xn, lw["qkv"], d)], rows, 3 * d, qkv, mid_dt, precision=prec)
            yield
            if decode:
                ops.timesfm_attention_decode(kv_cache[i], b, m.h, m.hd, patch_mask, num_masked, w["inv_freq"],
                                             lw["q_ln"], lw["k_ln"], lw["q_scale"], m.eps, adt, out=attn)
            else:
                ops.timesfm_attention(qkv, b, n, m.h, m.hd, patch_mask, num_masked, w["inv_freq"], lw["q_ln"],
                                      lw["k_ln"], lw["q_scale"], m.eps, adt, out=attn)
            yield
            if self.fused_norm:
                # out-proj / ff1 with post-norm + residual + next pre-norm in the GEMM epilogue (5-CTA clusters)
                ops.gemm_rownorm(attn, lw["out"], d, rows, d, prec, lw["post_attn"], lw["pre_ff"], x if i == 0 else y, y,
                                 adt, xn, m.eps)
                yield
                ops.gemm([(xn, lw["ff0"], d)], rows, m.ff, hbuf, adt, precision=prec, act=ACT_SILU)
                yield
                ops.gemm_rownorm(hbuf, lw["ff1"], m.ff, rows, d, prec, lw["post_ff"], nxt, y, y, adt,
                                 None if last else xn, m.eps)
                yield
            else:
                ops.gemm([(attn, lw["out"], d)], rows, d, a, mid_dt, precision=prec)
                yield
                ops.norm_residual_norm(a, x if i == 0 else y, lw["post_attn"], lw["pre_ff"], m.eps, y, adt, xn)
                yield
                ops.gemm([(xn, lw["ff0"], d)], rows, m.ff, hbuf, adt, precision=prec, act=ACT_SILU)
                yield
                ops.gemm([(hbuf, lw["ff1"], m.ff)], rows, d, a, mid_dt, precision=prec)
                yield
                ops.norm_residual_norm(a, y, lw["post_ff"], nxt, m.eps, y, adt, None if last else xn)
                yield
        return y

    # ------------------------------------------------------------------ beyond the reference: AR decode + extras
    def _head_steps(self, last: torch.Tensor, mu_last, sigma_last, out: torch.Tensor, head: str = "head"):
        """ResidualBlock head on [B, D] rows with the inverse RevIN (``acc * sigma + mu``) in the epilogue; ``out`` is a
        [B, width] fp32 view (row stride free) that receives all ``width`` columns."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        b, d = last.shape
        width = out.shape[1]
        a = ops.cast_rows(last, adt)
        yield
        hid = ops.alloc(b, m.md, adt, last.device)
        ops.gemm([(a, w[head + "_hidden"], d)], b, m.md, hid, adt, precision=prec, act=ACT_SILU)
        yield
        ops.gemm([(hid, w[head + "_out"], m.md), (a, w[head + "_res"], d)], b, width, out, DT_F32, precision=prec,
                 row_scale=sigma_last, row_shift=mu_last, ldd=out.stride(0))
        yield

    def decode_steps(self, horizon: int, inputs: torch.Tensor, masks: torch.Tensor, fuse=None):
        """Forecast (batch, horizon, 10) through upstream's decode loop, as a step generator: prefill on the context,
        ``(horizon - 1) // 128`` autoregressive steps that feed the previous 128-step point forecast back as 4 new
        patches (running RevIN statistics continued, attention against the KV cache the qkv GEMMs left in HBM), then
        the forecast extras of ``self.forecast_options`` in one finalize kernel.

        ``fuse``: generator function ``emb [B', N, D] -> emb`` applied to the context patches (text fusion; B' = 2 B
        under flip invariance, rows B.. being the negated series).  Raises the reference's ``ValueError`` for
        horizon > 128 unless ``forecast_options.ar_decode`` is set."""
        m, opts = self._model, self.forecast_options
        if horizon > m.o and not opts.ar_decode:
            raise ValueError(
                f"horizon must be <= output_patch_len ({m.o}), got {horizon}. AR decode is not supported."
            )
        steps = (horizon - 1) // m.o
        if steps + 1 > MAX_KV_REGIONS:
            raise ValueError(f"horizon {horizon} needs {steps} decode steps; at most {MAX_KV_REGIONS - 1} are supported")
        if opts.use_continuous_quantile_head and not hasattr(m, "output_projection_quantiles"):
            raise ValueError("use_continuous_quantile_head needs an adapter built with_quantile_head=True")
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        b = inputs.shape[0]
        flip = opts.force_flip_invariance
        masks = masks.bool()
        x_in = torch.cat([inputs, -inputs], dim=0) if flip else inputs
        m_in = torch.cat([masks, masks], dim=0) if flip else masks
        b2 = x_in.shape[0]
        pre = yield from self.preprocess_steps(x_in, m_in)
        emb = pre.input_embeddings
        if fuse is not None:
            emb = yield from fuse(emb)
        cache: list | None = [] if steps > 0 else None
        out_emb = yield from self.forward_steps(emb, pre.masks, kv_cache=cache)
        mu, sigma = pre.normalization_stats["context_mu"], pre.normalization_stats["context_sigma"]
        ht = m.o * (steps + 1)
        width = m.o * m.q
        pf = torch.empty(b2, ht * m.q, dtype=torch.float32, device=inputs.device)
        last = out_emb[:, -1, :]
        mu_last, sigma_last = mu[:, -1].contiguous(), sigma[:, -1].contiguous()
        yield from self._head_steps(last, mu_last, sigma_last, pf[:, :width])
        spread = None
        if opts.use_continuous_quantile_head:
            spread = torch.empty(b2, m.os * m.q, dtype=torch.float32, device=inputs.device)
            yield from self._head_steps(last, mu_last, sigma_last, spread, head="qhead")
            spread = spread.view(b2, m.os, m.q)
        if steps:
            per_step = m.o // m.p  # 4 new input patches per 128 forecast steps
            state = ((~m_in).sum(-1).float(), mu_last.clone(), sigma_last.clone())  # running (n, mu, sigma)
            patch_mask = pre.masks[..., -1].contiguous()
            num_masked = patch_mask.sum(-1, dtype=torch.int32)
            pf3 = pf.view(b2, ht, m.q)
            for s in range(steps):
                values = pf3[:, s * m.o : (s + 1) * m.o, m.config.decode_index]  # strided view, read in place
                tokens, mu_new, sigma_new = ops.timesfm_patchify_continue(values, state, per_step, m.p, adt)
                yield
                rows = b2 * per_step
                hidden = ops.alloc(rows, m.md, adt, inputs.device)
                ops.gemm([(tokens, w["tok_hidden"], 2 * m.p)], rows, m.md, hidden, adt, precision=prec, act=ACT_SILU,
                         bias=w["tok_hidden_b"])
                yield
                emb_new = torch.empty(rows, m.md, dtype=torch.float32, device=inputs.device)
                ops.gemm([(hidden, w["tok_out"], m.md), (tokens, w["tok_res"], 2 * m.p)], rows, m.md, emb_new, DT_F32,
                         precision=prec, bias=w["tok_out_b"])
                yield
                out_new = yield from self._stack_steps(emb_new, b2, per_step, patch_mask, num_masked, cache, decode=True)
                last_new = out_new.view(b2, per_step, m.md)[:, -1, :]
                yield from self._head_steps(last_new, mu_new[:, -1].contiguous(), sigma_new[:, -1].contiguous(),
                                            pf[:, (s + 1) * width : (s + 2) * width])
        out = ops.timesfm_forecast_finalize(
            pf.view(b2, ht, m.q), spread, inputs if opts.infer_is_positive else None, b, horizon,
            int(m.config.decode_index), flip, opts.use_continuous_quantile_head, opts.infer_is_positive,
        )
        yield
        return out

    def decode(self, horizon: int, inputs: torch.Tensor, masks: torch.Tensor, fuse=None) -> torch.Tensor:
        return drain(self.decode_steps(horizon, inputs, masks, fuse))

    def postprocess(
        self,
        horizon: int,
        output_embeddings: torch.Tensor,
        normalization_stats: dict[str, torch.Tensor],
    ) -> torch.Tensor:
        """Point/quantile head + inverse RevIN (reference timesfm.py:100-129) -> (batch, horizon, 10).

        Raises ValueError if ``horizon`` exceeds the output patch length (no AR decode).
        """
        return drain(self.postprocess_steps(horizon, output_embeddings, normalization_stats))

    def postprocess_steps(self, horizon: int, output_embeddings: torch.Tensor,
                          normalization_stats: dict[str, torch.Tensor]):
        """``postprocess`` as a step generator (see ``tsfmx_b200.lanes``)."""
        m = self._model
        if horizon > m.o:
            raise ValueError(
                f"horizon must be <= output_patch_len ({m.o}), got {horizon}. AR decode is not supported."
            )
        if not output_embeddings.is_cuda:
            raise TsfmxError("TimesFM2p5Adapter runs on B200 only; there is no CPU fallback")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        w = self._weights()
        b, n, d = output_embeddings.shape
        emb = output_embeddings.float()
        if emb.stride(-1) != 1 or emb.stride(1) != d:
            emb = emb.contiguous()
        last = emb[:, -1, :]  # only the last patch reaches the output (timesfm.py:129)
        mu_last = normalization_stats["context_mu"][:, -1].contiguous()
        sigma_last = normalization_stats["context_sigma"][:, -1].contiguous()
        a = ops.cast_rows(last, adt)
        yield
        hid = ops.alloc(b, m.md, adt, emb.device)
        ops.gemm([(a, w["head_hidden"], d)], b, m.md, hid, adt, precision=prec, act=ACT_SILU)
        yield
        out = torch.empty(b, horizon * m.q, dtype=torch.float32, device=emb.device)
        ops.gemm([(hid, w["head_out"], m.md), (a, w["head_res"], d)], b, m.o * m.q, out, DT_F32, precision=prec,
                 row_scale=sigma_last, row_shift=mu_last, n_store=horizon * m.q)
        yield
        return out.view(b, horizon, m.q)

    # ------------------------------------------------------------------ training path (frozen backbone)
    def forward_saving(self, input_embeddings: torch.Tensor, masks: torch.Tensor, for_wgrad: bool = False):
        """``forward`` that also keeps, per layer, what the activation-gradient pass needs: the layer input x,
        the mid-layer stream y, the raw qkv, the out-proj / ff1 outputs a1 / a2 and the ff0 pre-activation u.
        ``for_wgrad`` (full fine-tuning) additionally keeps the four GEMM inputs of every layer."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        b, n, d = input_embeddings.shape
        rows = b * n
        dev = input_embeddings.device
        x = input_embeddings.reshape(rows, d).float().contiguous()
        patch_mask = masks[..., -1].contiguous()
        num_masked = patch_mask.sum(-1, dtype=torch.int32)
        layers = w["layers"]
        saved = {"patch_mask": patch_mask, "num_masked": num_masked, "shape": (b, n, d), "layers": []}
        if not layers:
            return x.view(b, n, d), saved
        xn = ops.rmsnorm(x, layers[0]["pre_attn"], m.eps, adt)
        attn = ops.alloc(rows, d, adt, dev)
        hbuf = ops.alloc(rows, m.ff, adt, dev)
        cur = x
        for i, lw in enumerate(layers):
            last = i == len(layers) - 1
            xn1 = xn  # normed layer input (operand of the qkv GEMM)
            if for_wgrad:  # the weight-gradient GEMMs read these later: one set per layer instead of reused scratch
                attn = ops.alloc(rows, d, adt, dev)
                hbuf = ops.alloc(rows, m.ff, adt, dev)
                xn2 = ops.alloc(rows, d, adt, dev)
                nxt = None if last else ops.alloc(rows, d, adt, dev)
            else:
                xn2, nxt = xn, (None if last else xn)  # one scratch buffer, rewritten at both norm junctions
            qkv = ops.alloc(rows, 3 * d, mid_dt, dev)
            a1 = ops.alloc(rows, d, mid_dt, dev)
            a2 = ops.alloc(rows, d, mid_dt, dev)
            u = ops.alloc(rows, m.ff, mid_dt, dev)
            y = torch.empty(rows, d, dtype=torch.float32, device=dev)
            z = torch.empty(rows, d, dtype=torch.float32, device=dev)
            ops.gemm([(xn1, lw["qkv"], d)], rows, 3 * d, qkv, mid_dt, precision=prec)
            ops.timesfm_attention(qkv, b, n, m.h, m.hd, patch_mask, num_masked, w["inv_freq"], lw["q_ln"], lw["k_ln"],
                                  lw["q_scale"], m.eps, adt, out=attn)
            ops.gemm([(attn, lw["out"], d)], rows, d, a1, mid_dt, precision=prec)
            ops.norm_residual_norm(a1, cur, lw["post_attn"], lw["pre_ff"], m.eps, y, adt, xn2)
            ops.gemm([(xn2, lw["ff0"], d)], rows, m.ff, hbuf, adt, precision=prec, act=ACT_SILU, pre_act=u)
            ops.gemm([(hbuf, lw["ff1"], m.ff)], rows, d, a2, mid_dt, precision=prec)
            ops.norm_residual_norm(a2, y, lw["post_ff"], None if last else layers[i + 1]["pre_attn"], m.eps, z, adt, nxt)
            entry = {"x": cur, "y": y, "qkv": qkv, "a1": a1, "a2": a2, "u": u}
            if for_wgrad:
                entry.update({"xn1": xn1, "attn": attn, "xn2": xn2, "h": hbuf})
            saved["layers"].append(entry)
            cur = z
            xn = nxt
        return cur.view(b, n, d), saved

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, rows: int, n_out: int, k_in: int) -> torch.Tensor:
        """dW [n_out, k_in] = dY^T X as a K-major tcgen05 GEMM with K = rows (``ops.wgrad``)."""
        return ops.wgrad(dy, x, rows, n_out, k_in, PRECISIONS[self.precision])

    def forward_backward(self, saved, d_out: torch.Tensor, param_grads: dict[str, torch.Tensor] | None = None,
                         on_grads=None) -> torch.Tensor:
        """Activation gradient of ``forward``: dL/d(output embeddings) [M, D] fp32 -> dL/d(input embeddings).

        ``param_grads`` (full fine-tuning; needs ``forward_saving(for_wgrad=True)``): filled with the gradient of every
        parameter of the stack under its state-dict name.  ``on_grads(tensors)`` is called once per layer with that
        layer's finished parameter gradients (the trainer all-reduces them while the layers below are still running)."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w, wt = self._weights(), self._weights_t()
        b, n, d = saved["shape"]
        rows = b * n
        dev = d_out.device
        layers, tlayers, sl = w["layers"], wt["layers"], saved["layers"]
        if not layers:
            return d_out
        g = d_out.contiguous()  # running dL/dz (fp32), updated in place layer by layer
        g2 = ops.alloc(rows, d, adt, dev)  # gradient w.r.t. a GEMM output, as the next dgrad's A operand
        du = ops.alloc(rows, m.ff, adt, dev)
        gmid = ops.alloc(rows, d, mid_dt, dev)
        datt = ops.alloc(rows, d, mid_dt, dev)
        dqkv = ops.alloc(rows, 3 * d, adt, dev)
        pg = param_grads
        # full fine-tuning: the four norm-scale gradients of a layer leave the junction kernels' own pass over the rows
        # (views of one zero-filled buffer: the kernels accumulate into them)
        dscale = torch.zeros(len(layers), 4, d, dtype=torch.float32, device=dev) if pg is not None else None

        def slot(i: int, name: str):
            if dscale is None or i < 0:
                return None
            j = ("pre_attn_ln", "post_attn_ln", "pre_ff_ln", "post_ff_ln").index(name)
            pg[f"stacked_xf.{i}.{name}.scale"] = dscale[i, j]
            return dscale[i, j]

        # top of the stack: da2 = RMSNorm_bwd(a2, post_ff, dz)
        ops.rmsnorm_bwd_chain(g, None, None, None, sl[-1]["a2"], layers[-1]["post_ff"], m.eps, None, adt, g2, rows, d,
                              dw2=slot(len(layers) - 1, "post_ff_ln"))
        for i in reversed(range(len(layers))):
            lw, tw, s = layers[i], tlayers[i], sl[i]
            pre = f"stacked_xf.{i}."
            if pg is not None and on_grads is not None and i + 1 < len(layers):
                done = f"stacked_xf.{i + 1}."  # the layer above is complete: hand its gradients over
                on_grads([v for k, v in pg.items() if k.startswith(done)])
            if pg is not None:  # g = dL/dz, g2 = da2 at this point
                pg[pre + "ff1.weight"] = self._wgrad(g2, s["h"], rows, d, m.ff)
            # dhff = da2 W1 ; du = dhff * silu'(u)
            ops.gemm([(g2, tw["ff1"], d)], rows, m.ff, du, adt, precision=prec, act=ACT_SILU_GRAD, aux=s["u"])
            # dyn = du W0
            ops.gemm([(du, tw["ff0"], m.ff)], rows, d, gmid, mid_dt, precision=prec)
            if pg is not None:
                pg[pre + "ff0.weight"] = self._wgrad(du, s["xn2"], rows, m.ff, d)
            # dy = dz + RMSNorm_bwd(y, pre_ff, dyn) ; da1 = RMSNorm_bwd(a1, post_attn, dy)
            ops.rmsnorm_bwd_chain(g, s["y"], lw["pre_ff"], gmid, s["a1"], lw["post_attn"], m.eps, g, adt, g2, rows, d,
                                  dw1=slot(i, "pre_ff_ln"), dw2=slot(i, "post_attn_ln"))
            if pg is not None:  # g = dL/dy, g2 = da1
                pg[pre + "attn.out.weight"] = self._wgrad(g2, s["attn"], rows, d, d)
            # datt = da1 Wo
            ops.gemm([(g2, tw["out"], d)], rows, d, datt, mid_dt, precision=prec)
            dparams = torch.zeros(2 * m.hd, dtype=torch.float32, device=dev) if pg is not None else None
            ops.timesfm_attention_bwd(s["qkv"], datt, b, n, m.h, m.hd, saved["patch_mask"], saved["num_masked"],
                                      w["inv_freq"], lw["q_ln"], lw["k_ln"], lw["q_scale"], m.eps, adt, dqkv=dqkv,
                                      dparams=dparams)
            # dxn = dqkv Wqkv
            ops.gemm([(dqkv, tw["qkv"], 3 * d)], rows, d, gmid, mid_dt, precision=prec)
            if pg is not None:
                pg[pre + "attn.qkv_proj.weight"] = self._wgrad(dqkv, s["xn1"], rows, 3 * d, d)
                # q' = (q_ln * softplus(per_dim) * c) * qhat: chain rule for the three 80-vectors
                xf = m.stacked_xf[i].attn
                per_dim = xf.per_dim_scale.per_dim_scale.detach().float()
                c = 1.442695041 / math.sqrt(m.hd)
                d_eff, d_k = dparams[: m.hd], dparams[m.hd :]
                pg[pre + "attn.query_ln.scale"] = d_eff * lw["q_scale"]
                pg[pre + "attn.key_ln.scale"] = d_k.clone()
                pg[pre + "attn.per_dim_scale.per_dim_scale"] = d_eff * lw["q_ln"] * c * torch.sigmoid(per_dim)
            # dx = dy + RMSNorm_bwd(x, pre_attn, dxn) ; and, for the layer below, da2 = RMSNorm_bwd(a2, post_ff, dx)
            below = i - 1
            ops.rmsnorm_bwd_chain(g, s["x"], lw["pre_attn"], gmid, sl[below]["a2"] if below >= 0 else None,
                                  layers[below]["post_ff"] if below >= 0 else None, m.eps, g, adt,
                                  g2 if below >= 0 else None, rows, d,
                                  dw1=slot(i, "pre_attn_ln"), dw2=slot(below, "post_ff_ln"))
        if pg is not None and on_grads is not None:
            on_grads([v for k, v in pg.items() if k.startswith("stacked_xf.0.")])
        return g

    def preprocess_saving(self, inputs: torch.Tensor, masks: torch.Tensor):
        """``preprocess`` that keeps the tokenizer's operands (full fine-tuning)."""
        m = self._model
        batch_size, context = inputs.shape[0], inputs.shape[1]
        if context % m.p != 0:
            raise ValueError(f"context length ({context}) must be divisible by patch length ({m.p})")
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        n = context // m.p
        rows = batch_size * n
        masks = masks.bool()
        tokens, mu, sigma, _pm, _nm = ops.timesfm_patchify_norm(inputs, masks, m.p, adt)
        hidden = ops.alloc(rows, m.md, adt, inputs.device)
        z = ops.alloc(rows, m.md, mid_dt, inputs.device)
        ops.gemm([(tokens, w["tok_hidden"], 2 * m.p)], rows, m.md, hidden, adt, precision=prec, act=ACT_SILU,
                 bias=w["tok_hidden_b"], pre_act=z)
        emb = torch.empty(rows, m.md, dtype=torch.float32, device=inputs.device)
        ops.gemm([(hidden, w["tok_out"], m.md), (tokens, w["tok_res"], 2 * m.p)], rows, m.md, emb, DT_F32,
                 precision=prec, bias=w["tok_out_b"])
        pre = PreprocessResult(emb.view(batch_size, n, m.md), masks.reshape(batch_size, n, m.p),
                               {"context_mu": mu, "context_sigma": sigma})
        return pre, {"tokens": tokens, "hidden": hidden, "z": z, "rows": rows}

    def preprocess_backward(self, saved, d_emb: torch.Tensor, param_grads: dict[str, torch.Tensor]) -> None:
        """Weight / bias gradients of the tokenizer ResidualBlock from dL/d(input embeddings) [M, D] fp32."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        wt = self._weights_t()
        rows, tokens, hidden = saved["rows"], saved["tokens"], saved["hidden"]
        pg = param_grads
        pg["tokenizer.output_layer.weight"] = self._wgrad(d_emb, hidden, rows, m.md, m.md)
        pg["tokenizer.residual_layer.weight"] = self._wgrad(d_emb, tokens, rows, m.md, 2 * m.p)
        bias = ops.colsum_wgrad(d_emb)
        pg["tokenizer.output_layer.bias"] = bias
        pg["tokenizer.residual_layer.bias"] = bias.clone()
        d_emb_a = ops.cast_rows(d_emb, adt)
        dz32 = torch.empty(rows, m.md, dtype=torch.float32, device=d_emb.device)
        ops.gemm([(d_emb_a, wt["tok_out"], m.md)], rows, m.md, dz32, DT_F32, precision=prec, act=ACT_SILU_GRAD, aux=saved["z"])
        pg["tokenizer.hidden_layer.weight"] = self._wgrad(dz32, tokens, rows, m.md, 2 * m.p)
        pg["tokenizer.hidden_layer.bias"] = ops.colsum_wgrad(dz32)

    def postprocess_saving(self, horizon: int, output_embeddings: torch.Tensor, normalization_stats):
        """``postprocess`` that keeps the head's pre-activation and operands for the backward pass."""
        m = self._model
        if horizon > m.o:
            raise ValueError(
                f"horizon must be <= output_patch_len ({m.o}), got {horizon}. AR decode is not supported."
            )
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        b, n, d = output_embeddings.shape
        last = output_embeddings[:, -1, :]
        sigma_last = normalization_stats["context_sigma"][:, -1].contiguous()
        mu_last = normalization_stats["context_mu"][:, -1].contiguous()
        a = ops.cast_rows(last, adt)
        hid = ops.alloc(b, m.md, adt, last.device)
        z = ops.alloc(b, m.md, mid_dt, last.device)
        ops.gemm([(a, w["head_hidden"], d)], b, m.md, hid, adt, precision=prec, act=ACT_SILU, pre_act=z)
        out = torch.empty(b, horizon * m.q, dtype=torch.float32, device=last.device)
        ops.gemm([(hid, w["head_out"], m.md), (a, w["head_res"], d)], b, m.o * m.q, out, DT_F32, precision=prec,
                 row_scale=sigma_last, row_shift=mu_last, n_store=horizon * m.q)
        return out.view(b, horizon, m.q), {"z": z, "sigma": sigma_last, "horizon": horizon, "b": b, "n": n, "a": a,
                                           "hid": hid}

    def postprocess_backward(self, saved, grad_forecast: torch.Tensor,
                             param_grads: dict[str, torch.Tensor] | None = None) -> torch.Tensor:
        """dL/d(forecast) [B, h, 10] -> dL/d(output embeddings) [B, N, D] fp32 (non-zero in the last patch only);
        ``param_grads`` additionally receives the head's weight gradients (full fine-tuning)."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        wt = self._weights_t()
        b, horizon = saved["b"], saved["horizon"]
        dev = grad_forecast.device
        # undo the epilogue: out = acc * sigma + mu, only the first horizon*q columns exist
        dpre32 = torch.zeros(b, m.o * m.q, dtype=torch.float32, device=dev)
        dpre32[:, : horizon * m.q] = grad_forecast.reshape(b, horizon * m.q) * saved["sigma"][:, None]
        dpre = ops.cast_rows(dpre32, adt)
        dz = ops.alloc(b, m.md, adt, dev)
        ops.gemm([(dpre, wt["head_out"], m.o * m.q)], b, m.md, dz, adt, precision=prec, act=ACT_SILU_GRAD, aux=saved["z"])
        d_last = torch.empty(b, m.md, dtype=torch.float32, device=dev)
        ops.gemm([(dz, wt["head_hidden"], m.md), (dpre, wt["head_res"], m.o * m.q)], b, m.md, d_last, DT_F32,
                 precision=prec)
        if param_grads is not None:
            head = "output_projection_point."
            param_grads[head + "output_layer.weight"] = self._wgrad(dpre, saved["hid"], b, m.o * m.q, m.md)
            param_grads[head + "residual_layer.weight"] = self._wgrad(dpre, saved["a"], b, m.o * m.q, m.md)
            param_grads[head + "hidden_layer.weight"] = self._wgrad(dz, saved["a"], b, m.md, m.md)
        d_out = torch.zeros(b, saved["n"], m.md, dtype=torch.float32, device=dev)
        d_out[:, -1, :] = d_last  # only the last patch feeds the head (reference timesfm.py:129)
        return d_out

    # ------------------------------------------------------------------ checkpoints / freezing
    def load_checkpoint(self, path: str) -> None:
        """Load upstream TimesFM 2.5 safetensors (strict), reference timesfm.py:131-134."""
        from safetensors.torch import load_file

        self._model.load_state_dict(load_file(path), strict=True)

    @classmethod
    def from_pretrained(cls, device: torch.device, repo_id: str = "google/timesfm-2.5-200m-pytorch") -> "TimesFM2p5Adapter":
        """Download + load pretrained weights (reference timesfm.py:136-158); needs network access."""
        from huggingface_hub import hf_hub_download

        instance = cls()
        instance.to(device)
        instance.load_checkpoint(hf_hub_download(repo_id=repo_id, filename="model.safetensors"))
        return instance

    def freeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = False

    @property
    def graph_safe(self) -> bool:
        """The forecast path makes no host synchronisation and no allocation outside torch's allocator, so
        ``MultimodalDecoder`` may capture it into a CUDA graph."""
        return True

    def unfreeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = True

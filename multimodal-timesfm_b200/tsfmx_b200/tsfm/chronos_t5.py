"""Chronos-T5 adapter (BASELINE.json configs[2]: "Chronos-T5-base + text fusion, mean-scale/bin tokenisation +
encoder-decoder forecast"), B200-native.

Not part of the reference, which only wraps Chronos-2 (reference tsfmx/tsfm/chronos.py); provided as the extra plugin
SURVEY.md section 9 asks for, behind the reference's ``TsfmAdapter`` contract (reference tsfmx/tsfm/base.py:25-75):

  preprocess  : NaN-aware mean scaling + uniform-bin quantisation in one kernel (upstream
                ``chronos.MeanScaleUniformBins``; ids bit-exact), then the token-embedding gather
                -> (batch, context + 1, 768) with the EOS token appended; the fusion adds per-token text projections
  forward     : T5 encoder (12 blocks: RMS LayerNorm, fused QKV tcgen05 GEMM, tensor-core self-attention with the
                relative-position bias and no 1/sqrt(d), out-proj, ReLU MLP) -> (batch, context + 1, 768)
  postprocess : greedy autoregressive decoding of ``horizon`` tokens (per step and layer: cached self-attention,
                cross-attention over per-layer K/V of the encoder output projected once, ReLU MLP; tied LM head),
                EOS suppressed until ``horizon`` tokens exist (upstream passes min_new_tokens = prediction_length),
                then de-quantisation ``centers[id] * scale`` -> (batch, horizon, 1)

The parameter container uses the ``transformers`` T5ForConditionalGeneration key names (upstream ChronosModel.model),
so ``amazon/chronos-t5-*`` safetensors load unchanged.  Text embeddings are per TOKEN: ``expand_text_embeddings``
broadcasts per-patch embeddings (``text_patch_len`` steps each) to the context + EOS positions.
"""

from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops
from .._lib import ACT_RELU, DT_BF16, DT_F32, PREC_BF16, PRECISIONS, TsfmxError
from ..lanes import drain
from .base import PreprocessResult, TsfmAdapter


# --------------------------------------------------------------------------- tokenizer
class MeanScaleUniformBins(nn.Module):
    """chronos-t5-* tokenizer: n_tokens 4096, 2 special tokens (PAD 0, EOS 1), bin centres linspace(-15, 15, 4093)."""

    def __init__(self, n_tokens: int = 4096, n_special_tokens: int = 2, low_limit: float = -15.0,
                 high_limit: float = 15.0, pad_token_id: int = 0, eos_token_id: int = 1) -> None:
        super().__init__()
        self.n_tokens, self.n_special_tokens = n_tokens, n_special_tokens
        self.pad_token_id, self.eos_token_id = pad_token_id, eos_token_id
        centers = torch.linspace(low_limit, high_limit, n_tokens - n_special_tokens - 1)
        boundaries = torch.concat((torch.tensor([-1e20]), (centers[1:] + centers[:-1]) / 2, torch.tensor([1e20])))
        self.register_buffer("centers", centers, persistent=False)
        self.register_buffer("boundaries", boundaries, persistent=False)

    def context_input_transform(self, context: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """context (B, C) fp32, NaN = missing -> token ids (B, C+1) int64 with EOS, attention mask (B, C+1), scale (B,)."""
        return ops.chronos_t5_tokenize(context, self.boundaries, self.n_special_tokens, self.n_tokens,
                                       self.pad_token_id, self.eos_token_id)

    def output_transform(self, samples: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
        """token ids (B, L) -> values = centers[clamp(id - n_special - 1)] * scale."""
        return ops.chronos_t5_dequantize(samples, self.centers, scale, self.n_special_tokens)


# --------------------------------------------------------------------------- parameter containers (HF T5 key names)
class _T5Attention(nn.Module):
    def __init__(self, d_model: int, inner: int, heads: int, buckets: int, has_bias: bool) -> None:
        super().__init__()
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(buckets, heads)


class _T5Norm(nn.Module):
    def __init__(self, dims: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dims))


class _T5SelfLayer(nn.Module):
    def __init__(self, d_model, inner, heads, buckets, has_bias):
        super().__init__()
        self.SelfAttention = _T5Attention(d_model, inner, heads, buckets, has_bias)
        self.layer_norm = _T5Norm(d_model)


class _T5CrossLayer(nn.Module):
    def __init__(self, d_model, inner, heads, buckets):
        super().__init__()
        self.EncDecAttention = _T5Attention(d_model, inner, heads, buckets, False)
        self.layer_norm = _T5Norm(d_model)


class _T5Dense(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)


class _T5FFLayer(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.DenseReluDense = _T5Dense(d_model, d_ff)
        self.layer_norm = _T5Norm(d_model)


class _T5Block(nn.Module):
    def __init__(self, d_model, inner, heads, d_ff, buckets, has_bias, decoder):
        super().__init__()
        layers: list[nn.Module] = [_T5SelfLayer(d_model, inner, heads, buckets, has_bias)]
        if decoder:
            layers.append(_T5CrossLayer(d_model, inner, heads, buckets))
        layers.append(_T5FFLayer(d_model, d_ff))
        self.layer = nn.ModuleList(layers)


class _T5Stack(nn.Module):
    def __init__(self, num_layers, d_model, inner, heads, d_ff, buckets, decoder):
        super().__init__()
        self.block = nn.ModuleList(
            _T5Block(d_model, inner, heads, d_ff, buckets, i == 0, decoder) for i in range(num_layers)
        )
        self.final_layer_norm = _T5Norm(d_model)


class ChronosT5Module(nn.Module):
    """Parameters of a chronos-t5 backbone under the T5ForConditionalGeneration key names (defaults: t5-base shape,
    ReLU feed-forward, tied LM head, vocabulary 4096)."""

    def __init__(self, num_layers: int = 12, num_decoder_layers: int | None = None, d_model: int = 768, d_kv: int = 64,
                 num_heads: int = 12, d_ff: int = 3072, vocab_size: int = 4096, tie_word_embeddings: bool = True) -> None:
        super().__init__()
        if d_kv != 64:
            raise ValueError("the B200 attention kernels are built for d_kv = 64 (every chronos-t5 size uses it)")
        self.d_model, self.d_kv, self.num_heads, self.d_ff, self.vocab_size = d_model, d_kv, num_heads, d_ff, vocab_size
        self.inner = num_heads * d_kv
        self.num_buckets, self.max_distance, self.eps = 32, 128, 1e-6
        self.tie_word_embeddings = tie_word_embeddings
        self.pad_token_id, self.eos_token_id, self.decoder_start_token_id = 0, 1, 0
        dec_layers = num_layers if num_decoder_layers is None else num_decoder_layers
        self.shared = nn.Embedding(vocab_size, d_model)
        self.encoder = _T5Stack(num_layers, d_model, self.inner, num_heads, d_ff, self.num_buckets, False)
        self.decoder = _T5Stack(dec_layers, d_model, self.inner, num_heads, d_ff, self.num_buckets, True)
        if not tie_word_embeddings:
            self.lm_head = nn.Linear(d_model, vocab_size, bias=False)


def init_random_(model: nn.Module, seed: int = 0) -> None:
    """Deterministic random init of EVERY tensor, generated on the CPU: embeddings ~ N(0, 1) like T5, Linear ~ N(0, 0.03),
    norm weights 1 + 0.1 N(0, 1), relative-position bias ~ N(0, 0.5)."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("layer_norm.weight"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=gen)
            elif "relative_attention_bias" in name:
                v = 0.5 * torch.randn(p.shape, generator=gen)
            elif name.endswith("shared.weight"):
                v = torch.randn(p.shape, generator=gen)
            else:
                v = 0.03 * torch.randn(p.shape, generator=gen)
            p.copy_(v.to(p.device))


def relative_position_bucket(relative_position: torch.Tensor, bidirectional: bool, num_buckets: int = 32,
                             max_distance: int = 128) -> torch.Tensor:
    """T5's bucketing of (key position - query position); restated from the published algorithm (HF twin
    transformers/models/t5/modeling_t5.py ``T5Attention._relative_position_bucket``)."""
    buckets = torch.zeros_like(relative_position)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (relative_position > 0).to(torch.long) * num_buckets
        relative_position = relative_position.abs()
    else:
        relative_position = -torch.clamp(relative_position, max=0)
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.clamp(large, max=num_buckets - 1)
    return buckets + torch.where(is_small, relative_position, large)


# --------------------------------------------------------------------------- adapter
class ChronosT5Adapter(TsfmAdapter):
    """Chronos-T5 behind the three-stage adapter API; point forecast = greedy decoding (channel 0 of 1)."""

    text_patch_len = 32  # steps covered by one per-patch text embedding in ``expand_text_embeddings``

    def __init__(self, model: ChronosT5Module | None = None, precision: str = "bf16") -> None:
        super().__init__()
        self._model = model if model is not None else ChronosT5Module()
        self.tokenizer = MeanScaleUniformBins(self._model.vocab_size, 2, -15.0, 15.0, self._model.pad_token_id,
                                              self._model.eos_token_id)
        self.set_precision(precision)
        self._packed: dict[object, dict[str, object]] = {}
        self._bias_tables: dict[tuple, torch.Tensor] = {}
        self._graphs: dict[tuple, tuple] = {}
        self.use_cuda_graphs = True  # replay the greedy decoding loop from a captured graph
        # num_samples = 1: greedy decoding, one output channel.  num_samples > 1: upstream's probabilistic forecast -
        # that many sampled token paths per series (temperature / top-k as in ChronosConfig: 1.0 / 50), reduced to the
        # ``quantile_levels`` over the paths; the point forecast is the median channel.
        self.num_samples = 1
        self.temperature = 1.0
        self.top_k = 50
        self.quantile_levels = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)
        self.generator: torch.Generator | None = None  # optional CUDA generator for reproducible sampling

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision

    @property
    def model_dims(self) -> int:
        return int(self._model.d_model)

    @property
    def patch_len(self) -> int:
        return 1  # one embedding per time step (plus the EOS position)

    @property
    def num_outputs(self) -> int:
        """Channels of the ``postprocess`` output (not part of the reference contract; used for empty batches)."""
        return 1 if self.num_samples <= 1 else len(self.quantile_levels)

    @property
    def point_forecast_index(self) -> int:
        if self.num_samples <= 1:
            return 0
        return min(range(len(self.quantile_levels)), key=lambda i: abs(self.quantile_levels[i] - 0.5))

    def expand_text_embeddings(self, text_embeddings: torch.Tensor, context: int) -> torch.Tensor:
        """(batch, ceil(context / text_patch_len), E) per-patch text embeddings -> (batch, context + 1, E) per token;
        the EOS position gets a zero row (no text is added to it)."""
        steps = text_embeddings.shape[1] * self.text_patch_len
        if steps < context:
            raise ValueError(f"text_embeddings cover {steps} steps, context is {context}")
        per_token = text_embeddings.repeat_interleave(self.text_patch_len, dim=1)[:, steps - context:]
        eos = torch.zeros_like(per_token[:, :1])
        return torch.cat([per_token, eos], dim=1).contiguous()

    # ------------------------------------------------------------------ packed weights
    def _weights(self) -> dict[str, object]:
        prec = PRECISIONS[self.precision]
        params = list(self._model.parameters())
        key = (prec, tuple((p.data_ptr(), p._version) for p in params))
        if key in self._packed:
            return self._packed[key]
        self._packed.clear()
        self._bias_tables.clear()
        m = self._model
        adt = ops.act_dtype(prec)

        def pack(wt: torch.Tensor) -> torch.Tensor:
            return ops.cast_rows(wt.detach().float().contiguous(), adt)

        def f32(t: torch.Tensor) -> torch.Tensor:
            return t.detach().float().contiguous()

        def attn(a: _T5Attention) -> dict[str, torch.Tensor]:
            return {"qkv": pack(torch.cat([a.q.weight, a.k.weight, a.v.weight], 0)), "o": pack(a.o.weight)}

        w: dict[str, object] = {"shared": f32(m.shared.weight), "enc": [], "dec": []}
        for blk in m.encoder.block:
            sa, ff = blk.layer[0], blk.layer[1]
            w["enc"].append({**attn(sa.SelfAttention), "ln0": f32(sa.layer_norm.weight),
                             "wi": pack(ff.DenseReluDense.wi.weight), "wo": pack(ff.DenseReluDense.wo.weight),
                             "ln1": f32(ff.layer_norm.weight)})
        w["enc_final"] = f32(m.encoder.final_layer_norm.weight)
        w["enc_bias"] = f32(m.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight)  # [buckets, H]
        for blk in m.decoder.block:
            sa, ca, ff = blk.layer[0], blk.layer[1], blk.layer[2]
            c = ca.EncDecAttention
            w["dec"].append({**attn(sa.SelfAttention), "ln0": f32(sa.layer_norm.weight),
                             "cq": pack(c.q.weight), "ckv": pack(torch.cat([c.k.weight, c.v.weight], 0)),
                             "co": pack(c.o.weight), "ln1": f32(ca.layer_norm.weight),
                             "wi": pack(ff.DenseReluDense.wi.weight), "wo": pack(ff.DenseReluDense.wo.weight),
                             "ln2": f32(ff.layer_norm.weight)})
        w["dec_final"] = f32(m.decoder.final_layer_norm.weight)
        w["dec_bias"] = f32(m.decoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight)
        if m.tie_word_embeddings:  # HF rescales the decoder output by d_model ** -0.5 before the tied head
            w["lm_head"] = pack(m.shared.weight.detach().float() * (m.d_model ** -0.5))
        else:
            w["lm_head"] = pack(m.lm_head.weight)
        self._packed[key] = w
        return w

    def _bias_table(self, which: str, length: int, w: dict) -> tuple[torch.Tensor, int]:
        """fp32 [H, n] bias by (key - query): encoder n = 2 T - 1 (zero at T - 1), decoder n = L (zero at L - 1)."""
        emb = w["enc_bias" if which == "enc" else "dec_bias"]
        key = (which, length, emb.data_ptr())
        if key not in self._bias_tables:
            dev = emb.device
            if which == "enc":
                delta = torch.arange(-(length - 1), length, device=dev)
            else:
                delta = torch.arange(-(length - 1), 1, device=dev)
            bucket = relative_position_bucket(delta, which == "enc", self._model.num_buckets, self._model.max_distance)
            self._bias_tables[key] = emb[bucket].t().contiguous()
        return self._bias_tables[key], length - 1

    # ------------------------------------------------------------------ stages
    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> PreprocessResult:
        """Tokenise and embed.  ``masks`` True = padded (the tsfmx convention): padded steps become missing values
        (NaN -> PAD token, attention off).  Raises ValueError if the mask shape differs."""
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        if not inputs.is_cuda:
            raise TsfmxError("ChronosT5Adapter runs on B200 only; there is no CPU fallback")
        b, c = inputs.shape
        w = self._weights()
        x = torch.where(masks.bool(), torch.full_like(inputs, float("nan")), inputs.float())
        ids, attention_mask, scale = self.tokenizer.context_input_transform(x)
        emb = ops.embed_rows(ids, w["shared"])
        return PreprocessResult(
            input_embeddings=emb.view(b, c + 1, self._model.d_model),
            masks=~attention_mask,  # True = padded / missing; the EOS position is always attendable
            normalization_stats={"scale": scale, "token_ids": ids},
        )

    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """T5 encoder over the (fused) token embeddings -> (batch, tokens, d_model) encoder states."""
        return drain(self.forward_steps(input_embeddings, masks))

    def forward_steps(self, input_embeddings: torch.Tensor, masks: torch.Tensor):
        m = self._model
        if not input_embeddings.is_cuda:
            raise TsfmxError("ChronosT5Adapter runs on B200 only; there is no CPU fallback")
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        b, t, d = input_embeddings.shape
        rows = b * t
        dev = input_embeddings.device
        key_mask = (~masks.bool()).contiguous()
        bias, _ = self._bias_table("enc", t, w)
        x = input_embeddings.reshape(rows, d).float().contiguous().clone()  # residual stream, updated in place
        layers = w["enc"]
        xn = ops.rmsnorm(x, layers[0]["ln0"], m.eps, adt)
        yield
        qkv = ops.alloc(rows, 3 * m.inner, mid_dt, dev)
        attn = ops.alloc(rows, m.inner, adt, dev)
        a = ops.alloc(rows, d, mid_dt, dev)
        u = ops.alloc(rows, m.d_ff, adt, dev)
        final = torch.empty(rows, d, dtype=torch.float32, device=dev)
        for i, lw in enumerate(layers):
            ops.gemm([(xn, lw["qkv"], d)], rows, 3 * m.inner, qkv, mid_dt, precision=prec)
            yield
            ops.t5_encoder_attention(qkv, b, t, m.num_heads, key_mask, bias, adt, out=attn)
            yield
            ops.gemm([(attn, lw["o"], m.inner)], rows, d, a, mid_dt, precision=prec)
            yield
            ops.norm_residual_norm(a, x, None, lw["ln1"], m.eps, x, adt, xn)
            yield
            ops.gemm([(xn, lw["wi"], d)], rows, m.d_ff, u, adt, precision=prec, act=ACT_RELU)
            yield
            ops.gemm([(u, lw["wo"], m.d_ff)], rows, d, a, mid_dt, precision=prec)
            yield
            if i + 1 < len(layers):
                ops.norm_residual_norm(a, x, None, layers[i + 1]["ln0"], m.eps, x, adt, xn)
            else:
                ops.norm_residual_norm(a, x, None, w["enc_final"], m.eps, x, DT_F32, final)
            yield
        return final.view(b, t, d)

    def postprocess(
        self,
        horizon: int,
        output_embeddings: torch.Tensor,
        normalization_stats: dict[str, torch.Tensor],
    ) -> torch.Tensor:
        """Greedy decoding of ``horizon`` tokens + de-quantisation -> (batch, horizon, 1)."""
        ids = normalization_stats["token_ids"]
        attention_mask = ids != self._model.pad_token_id
        scale = normalization_stats["scale"]
        if self.num_samples <= 1:
            tokens, _ = self.decode(output_embeddings, attention_mask, horizon)
            return self.tokenizer.output_transform(tokens, scale).unsqueeze(-1)
        b, s = output_embeddings.shape[0], self.num_samples
        tokens, _ = self._decode_eager(output_embeddings, attention_mask, horizon, None, False, num_samples=s)
        values = self.tokenizer.output_transform(tokens, scale.repeat_interleave(s))  # [B * S, horizon]
        paths = values.view(b, s, horizon)
        q = torch.tensor(self.quantile_levels, dtype=torch.float32, device=paths.device)
        return torch.quantile(paths, q, dim=1).permute(1, 2, 0).contiguous()  # (batch, horizon, quantiles)

    def decode(self, encoder_states: torch.Tensor, attention_mask: torch.Tensor, horizon: int,
               forced_ids: torch.Tensor | None = None, return_logits: bool = False):
        """Autoregressive decoding over the encoder states.  Greedy by default; ``forced_ids`` (batch, horizon) teacher
        forces the inputs of steps 1.. (tests).  Returns (token ids (batch, horizon) int64, logits or None).

        Plain greedy decoding is launch bound (135 kernels per generated token), so the whole loop is captured once
        per (batch, tokens, horizon, precision) in a CUDA graph and replayed on later calls (``use_cuda_graphs``)."""
        if horizon < 1:
            raise ValueError(f"horizon must be >= 1, got {horizon}")
        if not encoder_states.is_cuda:
            raise TsfmxError("ChronosT5Adapter runs on B200 only; there is no CPU fallback")
        if forced_ids is not None or return_logits or not self.use_cuda_graphs or horizon < 4 or torch.cuda.is_current_stream_capturing():
            return self._decode_eager(encoder_states, attention_mask, horizon, forced_ids, return_logits)
        self._weights()  # pack (if needed) outside the capture
        params = tuple((q.data_ptr(), q._version) for q in self._model.parameters())
        # one graph (and one set of static buffers) per launching stream: two series lanes of the same shape must not
        # share them, their replays overlap in time
        key = (tuple(encoder_states.shape), horizon, self.precision, encoder_states.device.index,
               torch.cuda.current_stream().cuda_stream, params)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 4:  # each graph pins its KV caches: keep a few shapes at most
                self._graphs.clear()
            static_enc = encoder_states.float().contiguous().clone()
            static_mask = attention_mask.bool().contiguous().clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up run: sets function attributes, fills every lazy cache
                self._decode_eager(static_enc, static_mask, horizon, None, False)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                tokens, _ = self._decode_eager(static_enc, static_mask, horizon, None, False)
            # the packed weights the capture read are kept alive with it (a precision round trip clears self._packed while
            # this key can match again)
            entry = (graph, static_enc, static_mask, tokens, dict(self._packed))
            self._graphs[key] = entry
        graph, static_enc, static_mask, tokens, _packed = entry
        static_enc.copy_(encoder_states)
        static_mask.copy_(attention_mask.bool())
        graph.replay()
        return tokens.clone(), None

    def _decode_eager(self, encoder_states: torch.Tensor, attention_mask: torch.Tensor, horizon: int,
                      forced_ids: torch.Tensor | None, return_logits: bool, num_samples: int = 1):
        """``num_samples`` > 1: every series decodes that many sampled paths (rows b * S + s); the encoder-side keys /
        values are projected once per series and shared by its paths."""
        m = self._model
        prec = PRECISIONS[self.precision]
        adt = ops.act_dtype(prec)
        mid_dt = DT_BF16 if prec == PREC_BF16 else DT_F32
        w = self._weights()
        b, t, d = encoder_states.shape
        rows = b * t
        dev = encoder_states.device
        inner, heads, length = m.inner, m.num_heads, horizon
        key_mask = attention_mask.bool().contiguous()
        enc = ops.cast_rows(encoder_states.reshape(rows, d).float().contiguous(), adt)
        layers = w["dec"]
        # cross-attention keys / values of every layer: one GEMM each over all encoder positions
        cross = []
        for lw in layers:
            kv = ops.alloc(rows, 2 * inner, mid_dt, dev)
            ops.gemm([(enc, lw["ckv"], d)], rows, 2 * inner, kv, mid_dt, precision=prec)
            cross.append(kv)
        del enc
        samples = max(1, int(num_samples))
        b = b * samples  # decode rows: series-major, sample-minor (encoder-side tensors keep the series batch)
        caches = [ops.alloc(b * length, 3 * inner, mid_dt, dev).view(b, length, 3 * inner) for _ in layers]
        bias, bias_zero = self._bias_table("dec", length, w)
        attn = ops.alloc(b, inner, adt, dev)
        a = ops.alloc(b, d, mid_dt, dev)
        cq = ops.alloc(b, inner, mid_dt, dev)
        u = ops.alloc(b, m.d_ff, adt, dev)
        logits = torch.empty(b, m.vocab_size, dtype=torch.float32, device=dev)
        all_logits = torch.empty(b, length, m.vocab_size, dtype=torch.float32, device=dev) if return_logits else None
        tokens = torch.empty(b, length, dtype=torch.int64, device=dev)
        cur = torch.full((b,), m.decoder_start_token_id, dtype=torch.int64, device=dev)
        ld = 3 * inner
        for step in range(length):
            x = ops.embed_rows(cur, w["shared"])
            xn = ops.rmsnorm(x, layers[0]["ln0"], m.eps, adt)
            for i, lw in enumerate(layers):
                cache = caches[i]
                slot = cache[:, step, :]
                ops.gemm([(xn, lw["qkv"], d)], b, ld, slot, mid_dt, precision=prec)
                ops.t5_attention(slot, cache[:, :, inner:], cache[:, :, 2 * inner:], b, 1, step + 1, heads, adt, attn,
                                 q_rows=(ld, length * ld), kv_rows=(ld, length * ld), out_rows=(inner, inner),
                                 q_pos0=step, causal=True, bias=bias, bias_zero=bias_zero)
                ops.gemm([(attn, lw["o"], inner)], b, d, a, mid_dt, precision=prec)
                ops.norm_residual_norm(a, x, None, lw["ln1"], m.eps, x, adt, xn)
                ops.gemm([(xn, lw["cq"], d)], b, inner, cq, mid_dt, precision=prec)
                kv = cross[i]
                ops.t5_attention(cq, kv, kv[:, inner:], b, 1, t, heads, adt, attn, q_rows=(inner, inner),
                                 kv_rows=(2 * inner, t * 2 * inner), out_rows=(inner, inner), key_mask=key_mask,
                                 kv_batch_div=samples)
                ops.gemm([(attn, lw["co"], inner)], b, d, a, mid_dt, precision=prec)
                ops.norm_residual_norm(a, x, None, lw["ln2"], m.eps, x, adt, xn)
                ops.gemm([(xn, lw["wi"], d)], b, m.d_ff, u, adt, precision=prec, act=ACT_RELU)
                ops.gemm([(u, lw["wo"], m.d_ff)], b, d, a, mid_dt, precision=prec)
                nxt = layers[i + 1]["ln0"] if i + 1 < len(layers) else w["dec_final"]
                ops.norm_residual_norm(a, x, None, nxt, m.eps, x, adt, xn)
            ops.gemm([(xn, w["lm_head"], d)], b, m.vocab_size, logits, DT_F32, precision=prec)
            if all_logits is not None:
                all_logits[:, step] = logits
            # min_new_tokens = horizon: EOS is banned until the horizon is full.  Greedy = top-1; otherwise temperature /
            # top-k sampling as upstream's generate(do_sample=True, top_k=50, temperature=1.0) - ban, top-k, softmax and the
            # draw in one kernel, fed one uniform number per path
            if samples == 1:
                cur = ops.t5_sample_topk(logits, m.eos_token_id, top_k=1)
            else:
                draws = torch.rand(b, device=dev, generator=self.generator)
                cur = ops.t5_sample_topk(logits, m.eos_token_id, top_k=int(self.top_k or 0),
                                         temperature=float(self.temperature), uniform=draws)
            tokens[:, step] = cur
            if forced_ids is not None:
                cur = forced_ids[:, step].to(torch.int64).contiguous()
        return tokens, all_logits

    # ------------------------------------------------------------------ checkpoints / freezing
    def load_checkpoint(self, path: str) -> None:
        """Load a T5ForConditionalGeneration state dict (safetensors); the tied ``*.embed_tokens`` / ``lm_head`` aliases
        of a tied checkpoint are dropped, everything else must match strictly."""
        from safetensors.torch import load_file

        state = load_file(path)
        for alias in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight"):
            state.pop(alias, None)
        if self._model.tie_word_embeddings:
            state.pop("lm_head.weight", None)
        self._model.load_state_dict(state, strict=True)

    @classmethod
    def from_pretrained(cls, device: torch.device, repo_id: str = "amazon/chronos-t5-base") -> "ChronosT5Adapter":
        """Download + load pretrained weights; needs network access."""
        from huggingface_hub import hf_hub_download

        instance = cls()
        instance.to(device)
        instance.load_checkpoint(hf_hub_download(repo_id=repo_id, filename="model.safetensors"))
        return instance

    def freeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = False

    def unfreeze_parameters(self) -> None:
        for param in self.parameters():
            param.requires_grad = True

"""Chronos-T5 forecast path on the CUDA side (tokenise -> embed -> fusion -> T5 encoder -> greedy decoding ->
de-quantise) against the oracle built on the installed transformers T5 (oracle/chronos_t5_model_oracle.py).

Bars: token ids of the context bit-exact; encoder states and teacher-forced decoder logits <= 1e-3 relative in the
fp32-accumulate parity mode ("bf16x3"), bf16 mode tolerance stated separately; greedy token ids identical to HF
``generate`` in parity mode.
"""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import chronos_t5_model_oracle as T  # noqa: E402  (checker only)
from oracle import timesfm_oracle as O  # noqa: E402
from tsfmx_b200 import ops  # noqa: E402
from tsfmx_b200._lib import DT_BF16, DT_F32  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.chronos_t5 import ChronosT5Adapter, ChronosT5Module, init_random_, relative_position_bucket  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-3
BF16_TOL = 4e-2


def rel_max(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def build(layers, seed=0, tied=True):
    adapter = ChronosT5Adapter(ChronosT5Module(num_layers=layers, tie_word_embeddings=tied))
    init_random_(adapter._model, seed=seed)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    oracle = T.oracle_from_product(dec)
    return dec.to(DEV).eval(), oracle


def batch(b, context, padded, seed=5):
    ctx, masks, text, _ = O.synthetic_batch(b, context, 8, seed=seed, padded=padded, patch_len=32)
    return ctx * 2 + 0.5, masks, text


def test_relative_position_bucket_matches_hf():
    from transformers.models.t5.modeling_t5 import T5Attention

    delta = torch.arange(-700, 701)
    for bidir in (True, False):
        ref = T5Attention._relative_position_bucket(delta, bidirectional=bidir, num_buckets=32, max_distance=128)
        assert torch.equal(relative_position_bucket(delta, bidir, 32, 128), ref)


@pytest.mark.parametrize("t", [65, 513, 130, 33])
def test_encoder_attention_kernels(t):
    """Tensor-core and SIMT T5 attention cores against fp64 torch (bias by relative position, key mask, all-masked)."""
    b, h, hd = 4, 12, 64
    gen = torch.Generator(device=DEV).manual_seed(t)
    qkv = torch.randn(b * t, 3 * h * hd, generator=gen, device=DEV)
    qkv[:, : 2 * h * hd] *= 0.35
    km = torch.ones(b, t, dtype=torch.bool, device=DEV)
    km[1, : t // 3] = False
    km[2, :] = False
    km[3, 3::2] = False
    table = torch.randn(h, 2 * t - 1, generator=gen, device=DEV)
    q, k, v = qkv.to(torch.bfloat16).double().reshape(b, t, 3, h, hd).permute(2, 0, 3, 1, 4)
    idx = torch.arange(t, device=DEV)
    bias = table.double()[:, (idx[None, :] - idx[:, None]) + t - 1]  # [h, q, k]
    s = q @ k.transpose(-1, -2) + bias[None] + ((~km)[:, None, None, :] * torch.finfo(torch.float32).min).double()
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, h * hd).float()
    got = ops.t5_encoder_attention(qkv.to(torch.bfloat16), b, t, h, km, table, DT_BF16)
    assert got.dtype == torch.bfloat16
    assert rel_max(got.float(), ref) < 1e-2
    got32 = ops.t5_encoder_attention(qkv.to(torch.bfloat16).float(), b, t, h, km, table, DT_F32)
    assert rel_max(got32, ref) < 2e-5


@pytest.mark.parametrize("padded", [False, True])
def test_preprocess_and_encoder_parity(padded):
    dec, oracle = build(2)
    dec.set_precision("bf16x3")
    ctx, masks, text = batch(4, 128, padded)
    text_tok = dec.adapter.expand_text_embeddings(text, 128)
    with torch.no_grad():
        ref = oracle.adapter.preprocess(ctx, masks)
        got = dec.adapter.preprocess(ctx.to(DEV), masks.to(DEV))
        assert torch.equal(got.normalization_stats["token_ids"].cpu(), ref.normalization_stats["token_ids"])  # bit-exact
        assert torch.equal(got.masks.cpu(), ref.masks)
        assert torch.equal(got.normalization_stats["scale"].cpu(), ref.normalization_stats["scale"])
        assert torch.equal(got.input_embeddings.cpu(), ref.input_embeddings)
        fused_ref = oracle.fusion(ref.input_embeddings, text_tok)
        enc_ref = oracle.adapter(fused_ref, ref.masks)
        enc = dec.adapter(fused_ref.to(DEV), ref.masks.to(DEV))
    assert rel_max(enc.cpu(), enc_ref) < FP32_TOL
    dec.set_precision("bf16")
    with torch.no_grad():
        enc16 = dec.adapter(fused_ref.to(DEV), ref.masks.to(DEV))
    with torch.no_grad():  # tolerance derived from the oracle's own bf16 evaluation of the same encoder
        twin = O.bf16_oracle(oracle, O.T5_BF16_OUTPUTS)
        cal = rel_max(twin.adapter(fused_ref, ref.masks), enc_ref)
    err16 = rel_max(enc16.cpu(), enc_ref)
    print(f"chronos-t5 encoder bf16: product {err16:.3e}, bf16 oracle {cal:.3e}, ratio {err16 / cal:.2f}")
    assert err16 < O.BF16_TOL_FACTOR * cal, (err16, cal)
    assert err16 < BF16_TOL


@pytest.mark.parametrize("layers,context,horizon,padded,tied", [(2, 128, 12, True, True), (3, 64, 24, False, False)])
def test_decoder_parity_and_greedy_ids(layers, context, horizon, padded, tied):
    dec, oracle = build(layers, seed=3, tied=tied)
    dec.set_precision("bf16x3")
    ctx, masks, text = batch(5, context, padded, seed=9)
    text_tok = dec.adapter.expand_text_embeddings(text, context)
    with torch.no_grad():
        pre = oracle.adapter.preprocess(ctx, masks)
        enc_ref = oracle.adapter(oracle.fusion(pre.input_embeddings, text_tok), pre.masks)
        am = pre.normalization_stats["token_ids"] != 0
        ref_tokens = oracle.adapter.decode(enc_ref, am, horizon)
        ref_logits = oracle.adapter.teacher_forced_logits(enc_ref, am, ref_tokens)
        # product decoder driven with the ORACLE's encoder states and tokens: logits of every step
        _, logits = dec.adapter.decode(enc_ref.to(DEV), am.to(DEV), horizon, forced_ids=ref_tokens.to(DEV),
                                       return_logits=True)
        assert rel_max(logits.cpu(), ref_logits) < FP32_TOL
        tokens, _ = dec.adapter.decode(enc_ref.to(DEV), am.to(DEV), horizon)
        assert torch.equal(tokens.cpu(), ref_tokens)  # greedy ids identical to HF generate
        # random teacher-forced tokens: every cache slot holds a different key / value
        rnd = torch.randint(2, 4096, (5, horizon), generator=torch.Generator().manual_seed(1))
        rnd_logits = oracle.adapter.teacher_forced_logits(enc_ref, am, rnd)
        _, got_rnd = dec.adapter.decode(enc_ref.to(DEV), am.to(DEV), horizon, forced_ids=rnd.to(DEV), return_logits=True)
        assert rel_max(got_rnd.cpu(), rnd_logits) < FP32_TOL
        # whole path through the public API
        ref_full = oracle.forward_full(horizon, ctx, masks, text_tok)
        got_full = dec.forward_full(horizon, ctx.to(DEV), masks.to(DEV), text_tok.to(DEV)).cpu()
    assert got_full.shape == ref_full.shape == (5, horizon, 1)
    agree = (got_full == ref_full).float().mean().item()
    assert agree > 0.9, agree  # one flipped near-tie token changes the rest of that series' path
    dec.set_precision("bf16")
    with torch.no_grad():
        _, logits16 = dec.adapter.decode(enc_ref.to(DEV), am.to(DEV), horizon, forced_ids=ref_tokens.to(DEV),
                                         return_logits=True)
    assert rel_max(logits16.cpu(), ref_logits) < BF16_TOL


def test_graph_replay_and_series_lanes_give_the_eager_tokens():
    """Greedy decoding replayed from CUDA graphs, with the batch cut into two series lanes on separate streams (each
    lane owns its graph and static buffers), must reproduce the eager single-stream tokens exactly."""
    dec, _ = build(2, seed=5, tied=False)
    dec.set_precision("bf16")
    ctx, masks, text = batch(288, 64, True, seed=13)   # 288 x 65 tokens >= 2 x 8192: the decoder really uses two lanes
    text_tok = dec.adapter.expand_text_embeddings(text, 64)
    ctx, masks, text_tok = ctx.to(DEV), masks.to(DEV), text_tok.to(DEV)
    try:
        with torch.no_grad():
            dec.lanes, dec.adapter.use_cuda_graphs = 1, False
            eager = dec.forward_full(12, ctx, masks, text_tok)
            dec.adapter.use_cuda_graphs = True
            graphed = dec.forward_full(12, ctx, masks, text_tok)
            dec.lanes = 2
            assert dec._lane_count(ctx) == 2
            first = dec.forward_full(12, ctx, masks, text_tok)    # captures one graph per lane stream
            again = dec.forward_full(12, ctx, masks, text_tok)    # replays both, overlapping in time
        torch.cuda.synchronize()
    finally:
        dec.lanes, dec.adapter.use_cuda_graphs = 2, True
    assert torch.equal(eager, graphed) and torch.equal(eager, first) and torch.equal(eager, again)


def test_sampled_decoding_paths_and_quantiles():
    """num_samples > 1: sampled token paths share the encoder-side keys / values of their series.  top_k = 1 makes
    sampling deterministic, so every path must equal the greedy path; with top_k = 50 the quantile channels are ordered
    and reproducible under a seeded generator."""
    dec, _ = build(2, seed=7, tied=False)
    dec.set_precision("bf16x3")
    ad = dec.adapter
    ctx, masks, text = batch(5, 64, True, seed=17)
    text_tok = ad.expand_text_embeddings(text, 64)
    ctx, masks, text_tok = ctx.to(DEV), masks.to(DEV), text_tok.to(DEV)
    try:
        with torch.no_grad():
            greedy = dec.forward_full(10, ctx, masks, text_tok)                      # (5, 10, 1)
            ad.num_samples, ad.top_k = 4, 1
            same = dec.forward_full(10, ctx, masks, text_tok)                        # (5, 10, 9): all paths = greedy
            assert same.shape == (5, 10, 9) and ad.point_forecast_index == 4
            assert torch.equal(same, greedy.expand(-1, -1, 9))
            ad.top_k = 50
            ad.generator = torch.Generator(device=DEV).manual_seed(123)
            first = dec.forward_full(10, ctx, masks, text_tok)
            ad.generator.manual_seed(123)
            again = dec.forward_full(10, ctx, masks, text_tok)
            point = dec(10, ctx, masks, text_tok)
    finally:
        ad.num_samples, ad.top_k, ad.generator = 1, 50, None
    assert torch.equal(first, again)
    assert bool((first[..., 1:] >= first[..., :-1]).all())                            # quantiles are ordered
    assert point.shape == (5, 10) and torch.isfinite(first).all()


def test_sample_topk_kernel():
    """tsfmx_t5_sample_topk: greedy = first maximum with the banned id excluded; sampled ids come from softmax over the
    top-k survivors (frequencies over 40 000 draws), never a banned or non-top-k id; u -> 1 picks the last survivor."""
    from tsfmx_b200 import ops

    gen = torch.Generator(device=DEV).manual_seed(0)
    logits = torch.randn(64, 4096, generator=gen, device=DEV) * 3
    logits[0, 1] = 100.0  # the banned id would win
    logits[1, 7] = logits[1, 99] = 50.0  # a tie: first index
    greedy = ops.t5_sample_topk(logits, banned_id=1, top_k=1)
    masked = logits.clone()
    masked[:, 1] = float("-inf")
    assert torch.equal(greedy, masked.argmax(-1)) and greedy[1].item() == 7
    # one row, many draws
    row = torch.randn(4096, generator=gen, device=DEV) * 2
    row[1] = 30.0
    draws = 40000
    u = torch.rand(draws, generator=gen, device=DEV)
    for top_k, temperature in ((5, 1.0), (50, 0.7), (0, 1.0)):
        ids = ops.t5_sample_topk(row.expand(draws, -1).contiguous(), banned_id=1, top_k=top_k, temperature=temperature, uniform=u)
        ref = row.clone()
        ref[1] = float("-inf")
        ref = ref / temperature
        k = top_k if top_k else 4096
        top_v, top_i = torch.topk(ref, k)
        probs = torch.softmax(top_v, -1)
        assert bool(torch.isin(ids, top_i).all()) and not bool((ids == 1).any())
        freq = torch.bincount(ids, minlength=4096)[top_i].float() / draws
        # binomial standard error of the largest probabilities is ~2.5e-3 at 40 000 draws
        assert (freq - probs)[:5].abs().max().item() < 1.5e-2, (top_k, freq[:5], probs[:5])
        assert abs(freq.sum().item() - 1.0) < 1e-6
    last = ops.t5_sample_topk(row[None].contiguous(), banned_id=1, top_k=5, uniform=torch.ones(1, device=DEV))
    top5 = torch.topk(torch.where(torch.arange(4096, device=DEV) == 1, float("-inf"), row), 5).indices
    assert last.item() == top5.max().item()  # inverse CDF in id order: the highest id among the survivors


@pytest.mark.parametrize("tk,samples", [(513, 1), (129, 3), (64, 1), (700, 2)])
def test_cross_attention_decode_kernel_matches_the_general_kernel(tk, samples):
    """The decode step's cross-attention (one query row per path, bf16 K / V, key mask, sample paths sharing their
    series' keys) through its specialised kernel and through the general tsfmx_t5_attention kernel (tuning hook 5)."""
    from tsfmx_b200 import ops
    from tsfmx_b200._lib import DT_BF16, DT_F32

    gen = torch.Generator(device=DEV).manual_seed(tk)
    series, heads, inner = 7, 12, 768
    b = series * samples
    q = (torch.randn(b, inner, generator=gen, device=DEV) * 0.5).to(torch.bfloat16)
    kv = (torch.randn(series * tk, 2 * inner, generator=gen, device=DEV) * 0.5).to(torch.bfloat16)
    km = torch.rand(series, tk, generator=gen, device=DEV) > 0.2
    km[1] = False  # every key masked: uniform weights
    km[2] = True
    outs = {}
    for general in (1, 0):
        ops._lib.check(ops._lib.load().tsfmx_tune(5, general))
        try:
            for dt in (DT_F32, DT_BF16):
                out = ops.alloc(b, inner, dt, q.device)
                ops.t5_attention(q, kv, kv[:, inner:], b, 1, tk, heads, dt, out, q_rows=(inner, inner),
                                 kv_rows=(2 * inner, tk * 2 * inner), out_rows=(inner, inner), key_mask=km, kv_batch_div=samples)
                outs[(general, dt)] = out.float()
        finally:
            ops._lib.check(ops._lib.load().tsfmx_tune(5, 0))
    ref = outs[(1, DT_F32)]
    assert (outs[(0, DT_F32)] - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert (outs[(0, DT_BF16)] - outs[(1, DT_BF16)]).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())

"""TimesFM autoregressive decode (horizon > 128) and forecast extras on the GPU against the oracle.

Oracle status (oracle/timesfm_oracle.py): the extras are pinned to HF ``TimesFm2_5ModelForPrediction`` (CPU test
tests/test_oracle_cpu.py::test_forecast_extras_match_hf); the AR loop restates upstream ``decode`` and is declared
"parity unpinned".  The oracle recomputes the whole sequence every step; the product decodes 4 new tokens against the
qkv matrices earlier launches left in HBM, so these tests also check the KV-cache path against plain recomputation.
"""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import timesfm_oracle as O  # noqa: E402  (checker only)
from tsfmx_b200 import ops  # noqa: E402
from tsfmx_b200._lib import DT_BF16, DT_F32  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import ForecastOptions, TimesFM2p5Adapter, init_random_  # noqa: E402

DEV = "cuda"


def build(layers=2, quantile_head=True, seed=0):
    adapter = TimesFM2p5Adapter(num_layers=layers, with_quantile_head=quantile_head)
    init_random_(adapter, seed=seed)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    oracle = O.oracle_from_product(dec)
    return dec.to(DEV).eval(), oracle


@pytest.fixture(scope="module")
def pair():
    return build()


# ------------------------------------------------------------------------------------------------ kernels
def test_patchify_continue_matches_running_stats():
    torch.manual_seed(0)
    b = 37
    ctx = torch.randn(b, 96) * 2 + 1
    masks = torch.zeros(b, 96, dtype=torch.bool)
    masks[3, :40] = True
    adapter = O.OracleTimesFM2p5Adapter(1)
    mu, sigma, state = adapter._running_stats(ctx.reshape(b, -1, 32), masks.reshape(b, -1, 32))
    forecast = torch.randn(b, 128, 10) * 3
    new = forecast[:, :, 5].reshape(b, 4, 32)
    zeros = torch.zeros_like(new, dtype=torch.bool)
    ref_mu, ref_sigma, ref_state = adapter._running_stats(new, zeros, state)
    ref_tokens = torch.cat([O.revin(new, ref_mu, ref_sigma), torch.zeros_like(new)], dim=-1).reshape(b * 4, 64)
    dev_state = tuple(t.clone().to(DEV) for t in state)
    values = forecast.to(DEV)[:, :, 5]  # strided view
    tokens, got_mu, got_sigma = ops.timesfm_patchify_continue(values, dev_state, 4, 32, DT_F32)
    assert (got_mu.cpu() - ref_mu).abs().max() < 2e-6 and (got_sigma.cpu() - ref_sigma).abs().max() < 2e-6
    assert (tokens.cpu() - ref_tokens).abs().max() < 1e-5
    for got, ref in zip(dev_state, ref_state):
        assert (got.cpu() - ref).abs().max() < 2e-6  # state advanced in place
    split = ops.timesfm_patchify_continue(values, tuple(t.clone().to(DEV) for t in state), 4, 32, 2)[0]
    assert (ops.split_to_float(split).cpu() - ref_tokens).abs().max() < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n_ctx,steps", [(16, 1), (16, 3), (37, 2), (64, 4)])
def test_decode_attention_equals_full_attention_rows(dtype, n_ctx, steps):
    """Queries of the newest 4 tokens against regions [ctx | 4 | 4 | ...] == the last 4 rows of the full causal attention
    over the concatenated sequence (fp32 SIMT kernel), left padding included."""
    torch.manual_seed(n_ctx + steps)
    b, h, hd = 9, 16, 80
    total = n_ctx + 4 * steps
    qkv = (torch.randn(b, total, 3 * h * hd, device=DEV) * 0.7).to(dtype)
    pm = torch.zeros(b, total, dtype=torch.bool, device=DEV)
    pm[1, :3] = True
    pm[2, : n_ctx - 1] = True
    nm = pm.sum(-1, dtype=torch.int32)
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))).to(DEV)
    qw, kw, qs = (torch.rand(hd, device=DEV) + 0.5 for _ in range(3))
    ops._lib.load().tsfmx_attention_force_simt(1)
    try:
        full = ops.timesfm_attention(qkv.reshape(b * total, -1), b, total, h, hd, pm, nm, inv_freq, qw, kw, qs, 1e-6, DT_F32)
    finally:
        ops._lib.load().tsfmx_attention_force_simt(0)
    regions = [qkv[:, :n_ctx].reshape(b * n_ctx, -1).contiguous()]
    for s in range(steps):
        regions.append(qkv[:, n_ctx + 4 * s : n_ctx + 4 * s + 4].reshape(b * 4, -1).contiguous())
    got = ops.timesfm_attention_decode(regions, b, h, hd, pm[:, :n_ctx].contiguous(), nm, inv_freq, qw, kw, qs, 1e-6, DT_F32)
    want = full.view(b, total, h * hd)[:, -4:].reshape(b * 4, h * hd)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
    got16 = ops.timesfm_attention_decode(regions, b, h, hd, pm[:, :n_ctx].contiguous(), nm, inv_freq, qw, kw, qs, 1e-6, DT_BF16)
    assert (got16.float() - want).abs().max().item() < 1e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("cq,positive", [(False, False), (True, False), (True, True)])
def test_forecast_finalize_matches_hf_extras(flip, cq, positive):
    torch.manual_seed(5)
    b, ht, hs, q, horizon = 11, 384, 1024, 10, 300
    pf, spread = torch.randn((2 if flip else 1) * b, ht, q), torch.randn((2 if flip else 1) * b, hs, q)
    inputs = torch.randn(b, 64)
    inputs[::2] = inputs[::2].abs()  # every other series is non-negative
    opts = O.ForecastOptions(cq, flip, positive)
    ref = O.apply_forecast_extras(pf[:b], spread[:b], pf[b:] if flip else None, spread[b:] if flip else None, inputs,
                                  horizon, opts)
    got = ops.timesfm_forecast_finalize(pf.to(DEV), spread.to(DEV) if cq else None, inputs.to(DEV) if positive else None,
                                        b, horizon, 5, flip, cq, positive)
    assert got.shape == (b, horizon, q)
    assert (got.cpu() - ref).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------------------------ the whole path
OPTION_SETS = [
    ForecastOptions(ar_decode=True),
    ForecastOptions(ar_decode=True, use_continuous_quantile_head=True, force_flip_invariance=True, infer_is_positive=True),
]


@pytest.mark.parametrize("options", OPTION_SETS, ids=["ar", "ar+extras"])
@pytest.mark.parametrize("horizon", [64, 128, 129, 256, 300])
@pytest.mark.parametrize("padded", [False, True])
def test_decode_matches_oracle(pair, options, horizon, padded):
    dec, oracle = pair
    dec.set_precision("bf16x3")
    dec.adapter.forecast_options = options
    ctx, masks, text, _ = O.synthetic_batch(6, 512, 128, padded=padded, seed=31)
    ctx[0] = ctx[0].abs()  # one non-negative series (positivity clamp)
    o_opts = O.ForecastOptions(options.use_continuous_quantile_head, options.force_flip_invariance, options.infer_is_positive)
    try:
        with torch.no_grad():
            ref = oracle.forecast(horizon, ctx, masks, text, o_opts)
            got = dec.forward_full(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
            got_pt = dec(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    finally:
        dec.adapter.forecast_options = ForecastOptions()
    assert got.shape == ref.shape == (6, horizon, 10)
    err = O.rel_max(got, ref)
    print(f"decode h={horizon} padded={padded}: rel_max={err:.3e}")
    assert err < 1e-3, err
    assert O.rel_l2(got, ref) < 1e-3
    assert torch.equal(got_pt, got[..., 5])


def test_decode_without_text_long_context_and_bf16(pair):
    dec, oracle = pair
    ctx, masks, _text, _ = O.synthetic_batch(5, 2048, 128, padded=True, seed=7)
    dec.adapter.forecast_options = ForecastOptions(ar_decode=True, force_flip_invariance=True)
    try:
        with torch.no_grad():
            ref = oracle.forecast(256, ctx, masks, None, O.ForecastOptions(False, True, False))
            dec.set_precision("bf16x3")
            got = dec.forward_full(256, ctx.to(DEV), masks.to(DEV), None).cpu()
            assert O.rel_max(got, ref) < 1e-3
            dec.set_precision("bf16")
            ref16 = O.bf16_oracle(oracle).forecast(256, ctx, masks, None, O.ForecastOptions(False, True, False))
            got16 = dec.forward_full(256, ctx.to(DEV), masks.to(DEV), None).cpu()
            cal = O.rel_max(ref16, ref)
            print(f"decode bf16: product {O.rel_max(got16, ref):.3e}, bf16 oracle {cal:.3e}")
            assert O.rel_max(got16, ref) < O.BF16_TOL_FACTOR * cal
    finally:
        dec.adapter.forecast_options = ForecastOptions()
        dec.set_precision("bf16x3")


def test_decode_lanes_and_graph_replay_are_bit_identical(pair):
    dec, _ = pair
    dec.set_precision("bf16")
    dec.adapter.forecast_options = ForecastOptions(ar_decode=True, use_continuous_quantile_head=True)
    ctx, masks, text, _ = O.synthetic_batch(1100, 512, 128, padded=True, seed=2)
    ctx, masks, text = ctx.to(DEV), masks.to(DEV), text.to(DEV)
    saved = dec.lanes
    try:
        with torch.no_grad():
            dec.lanes = 1
            one = dec.forward_full(256, ctx, masks, text)
            dec.lanes = 2
            two = dec.forward_full(256, ctx, masks, text)
            dec.graphs = True
            replay = dec.forward_full(256, ctx, masks, text).clone()
            replay2 = dec.forward_full(256, ctx, masks, text).clone()
    finally:
        dec.lanes, dec.graphs = saved, False
        dec._graph_cache.clear()
        dec.adapter.forecast_options = ForecastOptions()
    assert torch.equal(one, two) and torch.equal(one, replay) and torch.equal(replay, replay2)


def test_reference_behaviour_is_the_default(pair):
    """Options off: horizon > 128 raises the reference's ValueError (timesfm.py:116-119); and the quantile-head option
    on an adapter built without that head is refused before any device work."""
    dec, _ = pair
    x = torch.zeros(2, 512, device=DEV)
    m = torch.zeros(2, 512, dtype=torch.bool, device=DEV)
    assert dec.adapter.forecast_options == ForecastOptions()
    with pytest.raises(ValueError, match="AR decode is not supported"):
        dec.forward_full(256, x, m, None)
    bare, _ = build(1, quantile_head=False)
    bare.adapter.forecast_options = ForecastOptions(use_continuous_quantile_head=True)
    with pytest.raises(ValueError, match="with_quantile_head"):
        bare.forward_full(64, x, m, None)
    bare.adapter.forecast_options = ForecastOptions(ar_decode=True)
    with pytest.raises(ValueError, match="decode steps"):
        bare.forward_full(128 * 17, x, m, None)

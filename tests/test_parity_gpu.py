"""Parity of the CUDA path (through the adapter API and the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star):
  * patch masks: bit-exact;
  * forecasts (all quantile channels, and the point forecast): <= 1e-3 relative in the fp32-accumulate
    parity mode ("bf16x3"), where relative = max|y - y_ref| / max|y_ref| per batch, and relative L2;
  * "bf16" throughput mode: tolerance stated separately below (BF16_TOL), calibrated on the oracle.
"""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import timesfm_oracle as O  # noqa: E402  (checker only)
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-3  # north_star: within 1e-3 relative (fp32 accumulate)
BF16_TOL = 4e-2  # bf16 operands, fp32 accumulate: stated separately (measured: see DESIGN.md)


def rel_max(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def build(num_layers, fusion_layers=1, hidden=None, seed=0):
    adapter = TimesFM2p5Adapter(num_layers=num_layers, with_quantile_head=False)
    init_random_(adapter, seed=seed)
    torch.manual_seed(seed + 100)  # fusion: the reference's own Xavier-uniform init
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, fusion_layers, list(hidden or [])))
    oracle = O.oracle_from_product(dec)
    return dec.to(DEV).eval(), oracle


@pytest.fixture(scope="module")
def model2():
    return build(2)


@pytest.fixture(scope="module")
def model20():
    return build(20)


@pytest.mark.parametrize("padded", [False, True])
def test_preprocess_stage(model2, padded):
    dec, oracle = model2
    dec.set_precision("bf16x3")
    ctx, masks, _text, _ = O.synthetic_batch(8, 512, 128, padded=padded)
    with torch.no_grad():
        ref = oracle.adapter.preprocess(ctx, masks)
        got = dec.adapter.preprocess(ctx.to(DEV), masks.to(DEV))
    assert torch.equal(got.masks.cpu(), ref.masks)  # bit-exact (B, N, 32) patch masks
    assert torch.equal(got.masks[..., -1].cpu(), ref.masks[..., -1])
    for k in ("context_mu", "context_sigma"):
        assert (got.normalization_stats[k].cpu() - ref.normalization_stats[k]).abs().max().item() < 1e-6
    assert rel_max(got.input_embeddings.cpu(), ref.input_embeddings) < 1e-4


@pytest.mark.parametrize("padded", [False, True])
def test_fusion_and_stack_stages(model2, padded):
    dec, oracle = model2
    dec.set_precision("bf16x3")
    ctx, masks, text, _ = O.synthetic_batch(8, 512, 128, padded=padded)
    with torch.no_grad():
        ref_pre = oracle.adapter.preprocess(ctx, masks)
        ref_fused = oracle.fusion(ref_pre.input_embeddings, text)
        ref_out = oracle.adapter(ref_fused, ref_pre.masks)
        # drive each product stage with the ORACLE's input so errors do not compound across stages
        fused = dec.fusion(ref_pre.input_embeddings.to(DEV), text.to(DEV))
        out = dec.adapter(ref_fused.to(DEV), ref_pre.masks.to(DEV))
    assert rel_max(fused.cpu(), ref_fused) < 1e-4
    assert rel_max(out.cpu(), ref_out) < 5e-4


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("horizon", [128, 32, 7])
def test_forward_full_parity_fp32_mode(model20, padded, horizon):
    dec, oracle = model20
    dec.set_precision("bf16x3")
    ctx, masks, text, _ = O.synthetic_batch(8, 512, horizon, padded=padded)
    with torch.no_grad():
        ref = oracle.forward_full(horizon, ctx, masks, text)
        got = dec.forward_full(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
        ref_pt = oracle(horizon, ctx, masks, text)
        got_pt = dec(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    assert got.shape == ref.shape == (8, horizon, 10)
    assert rel_max(got, ref) < FP32_TOL, rel_max(got, ref)
    assert rel_l2(got, ref) < FP32_TOL
    assert got_pt.shape == (8, horizon)
    assert rel_max(got_pt, ref_pt) < FP32_TOL


def test_forward_full_no_text_and_long_context(model2):
    dec, oracle = model2
    dec.set_precision("bf16x3")
    ctx, masks, _text, _ = O.synthetic_batch(5, 2048, 128, padded=True, seed=7)
    with torch.no_grad():
        ref = oracle.forward_full(128, ctx, masks, None)
        got = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), None).cpu()
    assert rel_max(got, ref) < FP32_TOL


def test_forward_full_bf16_mode(model20):
    dec, oracle = model20
    dec.set_precision("bf16")
    ctx, masks, text, _ = O.synthetic_batch(8, 512, 128, padded=False)
    with torch.no_grad():
        ref = oracle.forward_full(128, ctx, masks, text)
        got = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    err = rel_max(got, ref)
    print(f"bf16 mode: rel_max={err:.3e} rel_l2={rel_l2(got, ref):.3e}")
    assert err < BF16_TOL, err


@pytest.mark.parametrize("layers,hidden", [(2, [512]), (3, [1024, 512])])
def test_multi_layer_fusion(layers, hidden):
    dec, oracle = build(1, layers, hidden, seed=3)
    dec.set_precision("bf16x3")
    ctx, masks, text, _ = O.synthetic_batch(6, 512, 64, seed=11)
    with torch.no_grad():
        ref = oracle.forward_full(64, ctx, masks, text)
        got = dec.forward_full(64, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    assert rel_max(got, ref) < FP32_TOL


def test_zero_fusion_weights_is_identity(model2):
    # property from SURVEY.md section 4: fusion with zero weights == no fusion (relu(0) = 0)
    dec, _ = build(1, seed=5)
    dec.set_precision("bf16x3")
    with torch.no_grad():
        for lin in dec.fusion.linears():
            lin.weight.zero_()
    ctx, masks, text, _ = O.synthetic_batch(4, 512, 128)
    with torch.no_grad():
        a = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), text.to(DEV))
        b = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), None)
    assert torch.equal(a, b)


def test_error_behaviour_matches_reference(model2):
    dec, oracle = model2
    x = torch.zeros(2, 500, device=DEV)
    m = torch.zeros(2, 500, dtype=torch.bool, device=DEV)
    with pytest.raises(ValueError, match="divisible by patch length"):
        dec.forward_full(128, x, m, None)
    with pytest.raises(ValueError, match="must match inputs shape"):
        dec.forward_full(128, x, m[:, :32], None)
    x = torch.zeros(2, 512, device=DEV)
    m = torch.zeros(2, 512, dtype=torch.bool, device=DEV)
    with pytest.raises(ValueError, match="AR decode is not supported"):
        dec.forward_full(129, x, m, None)
    with pytest.raises(ValueError, match="AR decode is not supported"):
        oracle.forward_full(129, x.cpu(), m.cpu(), None)


@pytest.mark.parametrize("lanes_n,batch", [(2, 512), (3, 771), (4, 1024)])
def test_series_lanes_bit_identical(model2, lanes_n, batch):
    """Cutting the batch into series lanes on separate streams (tsfmx_b200.lanes) must not change a single bit:
    every series is independent on the whole path, so the lanes only change which kernels overlap in time."""
    dec, _ = model2
    dec.set_precision("bf16")
    ctx, masks, text, _ = O.synthetic_batch(batch, 512, 128, padded=True)
    ctx, masks, text = ctx.to(DEV), masks.to(DEV), text.to(DEV)
    saved = dec.lanes
    try:
        with torch.no_grad():
            dec.lanes = 1
            one = dec.forward_full(128, ctx, masks, text)
            dec.lanes = lanes_n
            assert dec._lane_count(ctx) == min(lanes_n, batch * 16 // 8192)
            many = dec.forward_full(128, ctx, masks, text)
            again = dec.forward_full(128, ctx, masks, text)
        torch.cuda.synchronize()
    finally:
        dec.lanes = saved
    assert torch.equal(one, many) and torch.equal(many, again)

"""Edge cases of the forecast path the reference's eager code handles implicitly: single and odd batch sizes, horizon 1,
fully padded series, empty batches, batches that do not divide into equal lanes."""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import timesfm_oracle as O  # noqa: E402  (checker only)
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

DEV = "cuda"


def rel_max(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def pair():
    adapter = TimesFM2p5Adapter(num_layers=2, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    oracle = O.oracle_from_product(dec)
    dec = dec.to(DEV).eval()
    dec.set_precision("bf16x3")
    return dec, oracle


@pytest.mark.parametrize("batch,context,horizon", [(1, 512, 128), (3, 32, 1), (7, 96, 17), (2, 4096, 128), (1, 7168, 64)])
def test_small_and_odd_shapes(pair, batch, context, horizon):
    dec, oracle = pair
    ctx, masks, text, _ = O.synthetic_batch(batch, context, horizon, seed=batch)
    with torch.no_grad():
        ref = oracle.forward_full(horizon, ctx, masks, text)
        got = dec.forward_full(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    assert got.shape == ref.shape == (batch, horizon, 10)
    assert rel_max(got, ref) < 1e-3


def test_fully_padded_series_and_scattered_padding(pair):
    dec, oracle = pair
    ctx, masks, text, _ = O.synthetic_batch(5, 256, 64, seed=11)
    masks = masks.clone()
    masks[0] = True            # nothing observed at all
    masks[1, ::3] = True       # scattered padding inside patches
    masks[2, :255] = True      # a single observed point
    with torch.no_grad():
        ref_pre = oracle.adapter.preprocess(ctx, masks)
        got_pre = dec.adapter.preprocess(ctx.to(DEV), masks.to(DEV))
        ref = oracle.forward_full(64, ctx, masks, text)
        got = dec.forward_full(64, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    assert torch.equal(got_pre.masks.cpu(), ref_pre.masks)  # bit-exact patch masks
    assert torch.isfinite(got).all() and torch.isfinite(ref).all()
    assert rel_max(got, ref) < 1e-3


def test_empty_batch_returns_empty_forecast(pair):
    dec, _ = pair
    ctx = torch.zeros(0, 512, device=DEV)
    out = dec.forward_full(128, ctx, torch.zeros(0, 512, dtype=torch.bool, device=DEV), torch.zeros(0, 16, 384, device=DEV))
    assert out.shape == (0, 128, 10)
    assert dec(64, ctx, torch.zeros(0, 512, dtype=torch.bool, device=DEV), None).shape == (0, 64)


def test_uneven_lanes_match_single_stream(pair):
    dec, _ = pair
    dec.set_precision("bf16")
    ctx, masks, text, _ = O.synthetic_batch(1031, 512, 128, seed=4, padded=True)
    ctx, masks, text = ctx.to(DEV), masks.to(DEV), text.to(DEV)
    try:
        with torch.no_grad():
            dec.lanes = 1
            one = dec.forward_full(128, ctx, masks, text)
            dec.lanes = 2
            two = dec.forward_full(128, ctx, masks, text)
    finally:
        dec.lanes = 2
        dec.set_precision("bf16x3")
    assert torch.equal(one, two)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device():
    """A decoder built on cuda:1 in a process whose current device is cuda:0: every entry point must run on the device
    that holds its operands (stream, TMA descriptors, SM count), and mixing devices must raise instead of launching."""
    from tsfmx_b200 import ops
    from tsfmx_b200._lib import DT_BF16, TsfmxError

    assert torch.cuda.current_device() == 0
    adapter = TimesFM2p5Adapter(num_layers=2, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig())
    ctx, masks, text, _ = O.synthetic_batch(520, 512, 128, padded=True)  # enough tokens for two series lanes
    with torch.no_grad():
        dec0 = dec.to("cuda:0").eval()
        want = dec0.forward_full(128, ctx.to("cuda:0"), masks.to("cuda:0"), text.to("cuda:0")).cpu()
        dec1 = dec.to("cuda:1").eval()
        a = dec1.forward_full(128, ctx.to("cuda:1"), masks.to("cuda:1"), text.to("cuda:1"))
        b = dec1.forward_full(128, ctx.to("cuda:1"), masks.to("cuda:1"), text.to("cuda:1"))  # second call: lanes
        dec1.graphs = True
        c = dec1.forward_full(128, ctx.to("cuda:1"), masks.to("cuda:1"), text.to("cuda:1"))
    assert a.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(a.cpu(), want) and torch.equal(b.cpu(), want) and torch.equal(c.cpu(), want)
    x0 = torch.randn(64, 128, device="cuda:0")
    with pytest.raises(TsfmxError, match="different devices"):
        ops.cast_rows(x0, DT_BF16, out=torch.empty(64, 128, dtype=torch.bfloat16, device="cuda:1"))

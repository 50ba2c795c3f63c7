"""Forecasts from a LOADED checkpoint (synthetic upstream-layout safetensors, tests/test_checkpoint_cpu.py) equal the
oracle's forecasts from the same file: the path a user of ``from_pretrained`` / ``load_checkpoint`` takes."""

import pytest
import torch
from safetensors.torch import load_file, save_file

pytestmark = pytest.mark.gpu

from oracle import chronos2_oracle as C  # noqa: E402  (checker only)
from oracle import timesfm_oracle as O  # noqa: E402
from test_checkpoint_cpu import upstream_chronos2_state_dict, upstream_timesfm_state_dict  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module  # noqa: E402
from tsfmx_b200.tsfm.timesfm import ForecastOptions, TimesFM2p5Adapter  # noqa: E402

DEV = "cuda"


def test_timesfm_forecast_from_loaded_checkpoint(tmp_path):
    path = tmp_path / "model.safetensors"
    save_file(upstream_timesfm_state_dict(3, seed=5), str(path))
    adapter = TimesFM2p5Adapter(num_layers=3, precision="bf16x3")
    adapter.to(DEV)                      # reference order (timesfm.py:152-157): construct, move, then load
    adapter.load_checkpoint(str(path))
    torch.manual_seed(1)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(DEV).eval()
    o_adapter = O.OracleTimesFM2p5Adapter(3, with_quantile_head=True)
    o_adapter.load_upstream_state_dict(load_file(str(path)))
    oracle = O.OracleDecoder(o_adapter, 384, 1, [])
    with torch.no_grad():
        oracle.fusion.projection[0].weight.copy_(dec.fusion.linears()[0].weight.cpu())
    ctx, masks, text, _ = O.synthetic_batch(6, 512, 128, padded=True, seed=4)
    with torch.no_grad():
        ref = oracle.forward_full(128, ctx, masks, text)
        got = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
        assert O.rel_max(got, ref) < 1e-3
        # the quantile head of the file is used too (continuous quantile head, HF modeling_timesfm2_5.py:816-826)
        adapter.forecast_options = ForecastOptions(use_continuous_quantile_head=True)
        ref_q = oracle.forecast(128, ctx, masks, text, O.ForecastOptions(True, False, False))
        got_q = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
        assert O.rel_max(got_q, ref_q) < 1e-3
        assert not torch.equal(got_q, got)


def test_chronos2_forecast_from_loaded_checkpoint(tmp_path):
    path = tmp_path / "model.safetensors"
    save_file(upstream_chronos2_state_dict(2, seed=6), str(path))
    adapter = Chronos2Adapter(Chronos2Module(2), precision="bf16x3")
    adapter.to(DEV)
    adapter.load_checkpoint(str(path))
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(DEV).eval()
    o_adapter = C.OracleChronos2Adapter(C.Chronos2Model(C.Chronos2Config(num_layers=2)))
    o_adapter.load_upstream_state_dict(load_file(str(path)))
    oracle = O.OracleDecoder(o_adapter, 384, 1, [])
    ctx, masks, _t, _ = O.synthetic_batch(5, 512, 64, padded=True, seed=8, patch_len=16)
    with torch.no_grad():
        ref = oracle.forward_full(64, ctx * 3 + 1, masks, None)
        got = dec.forward_full(64, (ctx * 3 + 1).to(DEV), masks.to(DEV), None).cpu()
    assert got.shape == ref.shape == (5, 64, 21)
    assert O.rel_max(got, ref) < 1e-3

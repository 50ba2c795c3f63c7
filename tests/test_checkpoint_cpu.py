"""Backbone checkpoint loading (reference tsfmx/tsfm/timesfm.py:131-158, chronos.py:171-199):
``load_checkpoint`` takes upstream safetensors with the upstream key names, strictly.  No network here, so the files are
synthetic: every tensor of the documented upstream layout (fused ``qkv_proj``,
``stacked_xf.{i}``, ``encoder.block.{i}.layer.{0,1,2}``, ...) with random values.  The GPU half of the check - forecasts
from a loaded checkpoint equal the oracle's on the same file - is tests/test_checkpoint_gpu.py."""

import pytest
import torch
from safetensors.torch import load_file, save_file

from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter


def upstream_timesfm_state_dict(num_layers: int, seed: int = 0) -> dict[str, torch.Tensor]:
    """google/timesfm-2.5-200m-pytorch layout (20 layers in the real file)."""
    g = torch.Generator().manual_seed(seed)
    d, hd, p, o, os_, q = 1280, 80, 32, 128, 1024, 10

    def w(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    for name, (d_in, d_hid, d_out, bias) in {"tokenizer": (2 * p, d, d, True), "output_projection_point": (d, d, o * q, False),
                                             "output_projection_quantiles": (d, d, os_ * q, False)}.items():
        for lin, (a, b) in {"hidden_layer": (d_hid, d_in), "output_layer": (d_out, d_hid), "residual_layer": (d_out, d_in)}.items():
            sd[f"{name}.{lin}.weight"] = w(a, b)
            if bias:
                sd[f"{name}.{lin}.bias"] = w(a, std=0.01)
    for i in range(num_layers):
        pre = f"stacked_xf.{i}."
        for ln in ("pre_attn_ln", "post_attn_ln", "pre_ff_ln", "post_ff_ln"):
            sd[pre + ln + ".scale"] = 1 + w(d, std=0.1)
        sd[pre + "attn.qkv_proj.weight"] = w(3 * d, d)
        sd[pre + "attn.out.weight"] = w(d, d)
        sd[pre + "attn.query_ln.scale"] = 1 + w(hd, std=0.1)
        sd[pre + "attn.key_ln.scale"] = 1 + w(hd, std=0.1)
        sd[pre + "attn.per_dim_scale.per_dim_scale"] = w(hd, std=0.5)
        sd[pre + "ff0.weight"] = w(d, d)
        sd[pre + "ff1.weight"] = w(d, d)
    return sd


def upstream_chronos2_state_dict(num_layers: int, seed: int = 0) -> dict[str, torch.Tensor]:
    """amazon/chronos-2 layout (12 blocks in the real file)."""
    g = torch.Generator().manual_seed(seed)
    d, inner, ff, nq = 768, 768, 3072, 21

    def w(*shape, std=0.03):
        return torch.randn(*shape, generator=g) * std

    sd = {"shared.weight": w(2, d), "encoder.final_layer_norm.weight": 1 + w(d, std=0.1)}
    for name, (d_in, d_out) in {"input_patch_embedding": (48, d), "output_patch_embedding": (d, nq * 16)}.items():
        for lin, (a, b) in {"hidden_layer": (ff, d_in), "output_layer": (d_out, ff), "residual_layer": (d_out, d_in)}.items():
            sd[f"{name}.{lin}.weight"] = w(a, b)
            sd[f"{name}.{lin}.bias"] = w(a, std=0.01)
    for i in range(num_layers):
        pre = f"encoder.block.{i}.layer."
        for j in (0, 1):
            for proj, shape in {"q": (inner, d), "k": (inner, d), "v": (inner, d), "o": (d, inner)}.items():
                sd[f"{pre}{j}.self_attention.{proj}.weight"] = w(*shape)
            sd[f"{pre}{j}.layer_norm.weight"] = 1 + w(d, std=0.1)
        sd[pre + "2.mlp.wi.weight"] = w(ff, d)
        sd[pre + "2.mlp.wo.weight"] = w(d, ff)
        sd[pre + "2.layer_norm.weight"] = 1 + w(d, std=0.1)
    return sd


def test_timesfm_checkpoint_loads_strictly(tmp_path):
    sd = upstream_timesfm_state_dict(2)
    path = tmp_path / "model.safetensors"
    save_file(sd, str(path))
    adapter = TimesFM2p5Adapter(num_layers=2)  # as the reference constructs it: with the quantile head
    adapter.load_checkpoint(str(path))
    got = adapter._model.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    # the adapter's own state dict = the upstream keys under "_model." (what trainer checkpoints of baseline mode hold)
    assert {k for k in adapter.state_dict()} == {"_model." + k for k in sd}
    # strict: a missing tensor or a stray one is an error, as in the reference (load_state_dict(strict=True))
    bad = dict(sd)
    bad.pop("stacked_xf.1.attn.qkv_proj.weight")
    save_file(bad, str(path))
    with pytest.raises(RuntimeError, match="Missing key"):
        TimesFM2p5Adapter(num_layers=2).load_checkpoint(str(path))
    bad = dict(sd)
    bad["stacked_xf.0.attn.q_proj.weight"] = torch.zeros(1280, 1280)  # un-fused HF-style key: not the upstream layout
    save_file(bad, str(path))
    with pytest.raises(RuntimeError, match="Unexpected key"):
        TimesFM2p5Adapter(num_layers=2).load_checkpoint(str(path))
    # a checkpoint of another depth does not load into this one
    save_file(upstream_timesfm_state_dict(3), str(path))
    with pytest.raises(RuntimeError):
        TimesFM2p5Adapter(num_layers=2).load_checkpoint(str(path))


def test_chronos2_checkpoint_loads_strictly(tmp_path):
    sd = upstream_chronos2_state_dict(2)
    path = tmp_path / "model.safetensors"
    save_file(sd, str(path))
    adapter = Chronos2Adapter(Chronos2Module(2))
    adapter.load_checkpoint(str(path))
    got = adapter._model.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    bad = dict(sd)
    bad.pop("encoder.block.1.layer.1.self_attention.v.weight")
    save_file(bad, str(path))
    with pytest.raises(RuntimeError, match="Missing key"):
        Chronos2Adapter(Chronos2Module(2)).load_checkpoint(str(path))
    assert set(load_file(str(path))) == set(bad)

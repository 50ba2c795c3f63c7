"""CUDA path against the committed golden vectors (tests/golden/*.npz, produced by the reference's own decoder
class around the oracle adapters; see tests/golden/make_golden.py).  Nothing here touches /root/reference."""

from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tsfmx_b200 import ops  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
DEV = "cuda"


def load_case(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize(
    "name", ["timesfm_l2_b4_c512_h128", "timesfm_l20_b2_c512_h128", "timesfm_l2_b3_c2048_h64_f2"]
)
def test_timesfm_golden(name):
    g = load_case(name)
    seed = int(g["seed"])
    adapter = TimesFM2p5Adapter(num_layers=int(g["num_layers"]), with_quantile_head=False)
    init_random_(adapter, seed=seed)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, int(g["fusion_layers"]), g["hidden"].tolist()))
    dec = dec.to(DEV).eval()
    dec.set_precision("bf16x3")
    ctx = torch.from_numpy(g["context"]).to(DEV)
    masks = torch.from_numpy(g["masks"]).to(DEV)
    text = torch.from_numpy(g["text"]).float().to(DEV)
    h = int(g["horizon"])
    with torch.no_grad():
        pre = dec.adapter.preprocess(ctx, masks)
        full = dec.forward_full(h, ctx, masks, text).cpu().numpy()
        point = dec(h, ctx, masks, text).cpu().numpy()
        no_text = dec.forward_full(h, ctx, masks, None).cpu().numpy()
    assert np.array_equal(pre.masks[..., -1].cpu().numpy(), g["patch_mask"])  # bit-exact
    np.testing.assert_allclose(pre.normalization_stats["context_mu"].cpu().numpy(), g["context_mu"], atol=2e-6)
    np.testing.assert_allclose(pre.normalization_stats["context_sigma"].cpu().numpy(), g["context_sigma"], atol=2e-6)
    emb_sum = pre.input_embeddings.double().sum(-1).cpu().numpy()
    assert np.abs(emb_sum - g["emb_checksum"]).max() < 1e-3 * max(1.0, np.abs(g["emb_checksum"]).max())
    scale = np.abs(g["forecast"]).max()
    assert np.abs(full - g["forecast"]).max() < 1e-3 * scale
    assert np.abs(point - g["point"]).max() < 1e-3 * scale
    assert np.abs(no_text - g["forecast_no_text"]).max() < 1e-3 * scale


def test_chronos_t5_golden_bit_exact():
    from tsfmx_b200.tsfm.chronos_t5 import MeanScaleUniformBins

    g = load_case("chronos_t5_tokens")
    tok = MeanScaleUniformBins().to(DEV)
    ids, am, scale = tok.context_input_transform(torch.from_numpy(g["x"]).to(DEV))
    assert np.array_equal(ids.cpu().numpy(), g["ids"].astype(np.int64))
    assert np.array_equal(am.cpu().numpy(), g["attention_mask"])
    assert np.array_equal(scale.cpu().numpy(), g["scale"])


def test_chronos2_golden():
    from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module
    from tsfmx_b200.tsfm.chronos import init_random_ as c2_init

    g = load_case("chronos2_l2_b3_c500_h40")
    module = Chronos2Module(2)
    c2_init(module, 0)
    dec = MultimodalDecoder(Chronos2Adapter(module), MultimodalDecoderConfig(384, 1, []))
    with torch.no_grad():
        dec.fusion.linears()[0].weight.copy_(torch.from_numpy(g["fusion_weight"]))
    dec = dec.to(DEV).eval()
    dec.set_precision("bf16x3")
    ctx, masks = torch.from_numpy(g["context"]).to(DEV), torch.from_numpy(g["masks"]).to(DEV)
    with torch.no_grad():
        pre = dec.adapter.preprocess(ctx, masks)
        full = dec.forward_full(40, ctx, masks, torch.from_numpy(g["text"]).float().to(DEV)).cpu().numpy()
    assert np.array_equal(pre.masks.cpu().numpy(), g["patch_mask"])  # bit-exact
    np.testing.assert_allclose(pre.normalization_stats["loc"].cpu().numpy(), g["loc"], rtol=1e-5, atol=1e-6)
    assert np.abs(full - g["forecast"]).max() < 1e-3 * np.abs(g["forecast"]).max()



def test_chronos_t5_model_golden():
    """Chronos-T5 forecast against the fixture made by the reference's decoder class around transformers' T5."""
    from tsfmx_b200.tsfm.chronos_t5 import ChronosT5Adapter, ChronosT5Module
    from tsfmx_b200.tsfm.chronos_t5 import init_random_ as t5_init

    g = load_case("chronos_t5_model_l2_b4_c96_h16")
    adapter = ChronosT5Adapter(ChronosT5Module(num_layers=2, tie_word_embeddings=False))
    t5_init(adapter._model, 0)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    torch.manual_seed(100)
    torch.nn.init.xavier_uniform_(dec.fusion.linears()[0].weight)  # same seeded CPU init as make_golden.py
    dec = dec.to(DEV).eval()
    dec.set_precision("bf16x3")
    ctx, masks = torch.from_numpy(g["context"]).to(DEV), torch.from_numpy(g["masks"]).to(DEV)
    text = dec.adapter.expand_text_embeddings(torch.from_numpy(g["text"]).float(), 96).to(DEV)
    with torch.no_grad():
        pre = dec.adapter.preprocess(ctx, masks)
        enc = dec.adapter(dec.fusion(pre.input_embeddings, text), pre.masks)
        tokens, _ = dec.adapter.decode(enc, pre.normalization_stats["token_ids"] != 0, 16)
        full = dec.forward_full(16, ctx, masks, text).cpu().numpy()
    assert np.array_equal(pre.normalization_stats["token_ids"].cpu().numpy(), g["token_ids"].astype(np.int64))  # bit-exact
    assert np.array_equal(pre.normalization_stats["scale"].cpu().numpy(), g["scale"])
    chk = enc.double().sum(-1).cpu().numpy()
    assert np.abs(chk - g["encoder_checksum"]).max() < 1e-3 * max(1.0, np.abs(g["encoder_checksum"]).max())
    assert np.array_equal(tokens.cpu().numpy(), g["generated"].astype(np.int64))  # greedy ids identical
    assert np.array_equal(full, g["forecast"])  # same ids, same centers, same scale -> identical values

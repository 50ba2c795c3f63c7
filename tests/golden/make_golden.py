"""Generate the golden vectors under tests/golden/ (run in the build container, where /root/reference exists).

    python tests/golden/make_golden.py

The reference ships no fixtures for this path (its tests/ directory is empty), so the vectors are produced by
running the reference's OWN ``MultimodalDecoder`` + ``MultimodalFusion`` (imported from
/root/reference/src) around the oracle adapter (oracle/timesfm_oracle.py) on seeded inputs and seeded
random-init weights.  Weights are not stored: they are regenerated from the seed by
``tsfmx_b200.tsfm.timesfm.init_random_`` (CPU torch generator) and ``torch.manual_seed`` (fusion Xavier init).
"""

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "multimodal-timesfm_b200"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")

from oracle import chronos_t5_oracle as T5  # noqa: E402
from oracle import timesfm_oracle as O  # noqa: E402
from tsfmx.decoder import MultimodalDecoder as RefDecoder  # noqa: E402  (the reference's real class)
from tsfmx.decoder import MultimodalDecoderConfig as RefConfig  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

OUT = Path(__file__).resolve().parent


def timesfm_case(name, num_layers, batch, context, horizon, padded, fusion_layers=1, hidden=(), seed=0):
    adapter = TimesFM2p5Adapter(num_layers=num_layers, with_quantile_head=False)
    init_random_(adapter, seed=seed)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, fusion_layers, list(hidden)))
    oracle = O.oracle_from_product(dec)
    ref = RefDecoder(oracle.adapter, RefConfig(384, fusion_layers, list(hidden))).eval()
    ref.fusion.load_state_dict(oracle.fusion.state_dict())
    ctx, masks, text, _ = O.synthetic_batch(batch, context, horizon, padded=padded, seed=1234 + seed)
    text = text.half().float()  # stored as fp16: round first so both sides see identical inputs
    with torch.no_grad():
        pre = ref.adapter.preprocess(ctx, masks)
        full = ref.forward_full(horizon, ctx, masks, text)
        point = ref(horizon, ctx, masks, text)
        no_text = ref.forward_full(horizon, ctx, masks, None)
    np.savez_compressed(
        OUT / f"{name}.npz",
        num_layers=num_layers, fusion_layers=fusion_layers, hidden=np.array(hidden, dtype=np.int64), seed=seed,
        horizon=horizon, context=ctx.numpy(), masks=masks.numpy(), text=text.numpy().astype(np.float16),
        context_mu=pre.normalization_stats["context_mu"].numpy(),
        context_sigma=pre.normalization_stats["context_sigma"].numpy(),
        patch_mask=pre.masks[..., -1].numpy(),
        emb_checksum=pre.input_embeddings.double().sum(-1).numpy(),
        forecast=full.numpy(), point=point.numpy(), forecast_no_text=no_text.numpy(),
    )
    print(name, full.shape, float(full.abs().max()))


def t5_case():
    g = torch.Generator().manual_seed(77)
    x = torch.randn(16, 512, generator=g) * torch.rand(16, 1, generator=g) * 8
    x[0, :9] = float("nan")
    x[1] = float("nan")
    x[2] = 0.0
    x[3] *= 1e3
    x[5] *= 0.01
    x[5, 100], x[5, 200] = 1e6, -1e6  # outliers -> clamp to the outer bins
    centers, boundaries = T5.tables()
    x[4, :500] = boundaries[1800:2300]
    ids, am, scale = T5.tokenize(x, boundaries)
    np.savez_compressed(OUT / "chronos_t5_tokens.npz", x=x.numpy(), ids=ids.numpy().astype(np.int16),
                        attention_mask=am.numpy(), scale=scale.numpy())
    print("t5", ids.shape, int(ids.min()), int(ids.max()))


def chronos2_case():
    from oracle import chronos2_oracle as C
    from tsfmx_b200.tsfm.chronos import Chronos2Module
    from tsfmx_b200.tsfm.chronos import init_random_ as c2_init

    module = Chronos2Module(2)
    c2_init(module, 0)
    o_adapter = C.OracleChronos2Adapter(C.Chronos2Model(C.Chronos2Config(num_layers=2)))
    o_adapter.load_upstream_state_dict(module.state_dict())
    ref = RefDecoder(o_adapter, RefConfig(384, 1, [])).eval()  # the reference's real decoder + fusion classes
    torch.manual_seed(100)
    torch.nn.init.xavier_uniform_(ref.fusion.projection[0].weight)
    ctx, masks, _t, _ = O.synthetic_batch(3, 500, 40, padded=True, seed=9, patch_len=16)
    ctx = ctx * 2 + 0.5
    g = torch.Generator().manual_seed(9)
    text = torch.randn(3, 32, 384, generator=g)
    text = (text / text.norm(dim=-1, keepdim=True)).half().float()
    with torch.no_grad():
        pre = ref.adapter.preprocess(ctx, masks)
        full = ref.forward_full(40, ctx, masks, text)
    np.savez_compressed(
        OUT / "chronos2_l2_b3_c500_h40.npz", context=ctx.numpy(), masks=masks.numpy(), text=text.numpy().astype(np.float16),
        fusion_weight=ref.fusion.projection[0].weight.detach().numpy().astype(np.float32),
        patch_mask=pre.masks.numpy(), loc=pre.normalization_stats["loc"].numpy(),
        scale=pre.normalization_stats["scale"].numpy(), forecast=full.numpy(),
    )
    print("chronos2", full.shape, float(full.abs().max()))


def chronos_t5_model_case():
    """Chronos-T5 forecast (2 + 2 layers of the t5-base shape): the reference's real decoder / fusion classes around
    the oracle adapter, i.e. transformers' T5ForConditionalGeneration with the seeded weights of the product module."""
    from oracle import chronos_t5_model_oracle as TM
    from tsfmx_b200.tsfm.chronos_t5 import ChronosT5Adapter, ChronosT5Module
    from tsfmx_b200.tsfm.chronos_t5 import init_random_ as t5_init

    adapter = ChronosT5Adapter(ChronosT5Module(num_layers=2, tie_word_embeddings=False))  # untied head: varied tokens
    t5_init(adapter._model, 0)
    o_adapter = TM.OracleChronosT5Adapter(TM.hf_model_from_product(adapter))
    ref = RefDecoder(o_adapter, RefConfig(384, 1, [])).eval()
    torch.manual_seed(100)
    torch.nn.init.xavier_uniform_(ref.fusion.projection[0].weight)
    ctx, masks, text, _ = O.synthetic_batch(4, 96, 16, padded=True, seed=21, patch_len=32)
    ctx = ctx * 2 + 0.5
    patch_text = text.half().float()  # stored per patch (fp16); the tests expand it the same way
    text = adapter.expand_text_embeddings(patch_text, 96)  # per-token text rows (context + EOS)
    horizon = 16
    with torch.no_grad():
        pre = ref.adapter.preprocess(ctx, masks)
        enc = ref.adapter(ref.fusion(pre.input_embeddings, text), pre.masks)
        tokens = ref.adapter.decode(enc, pre.normalization_stats["token_ids"] != 0, horizon)
        full = ref.forward_full(horizon, ctx, masks, text)
    np.savez_compressed(
        OUT / "chronos_t5_model_l2_b4_c96_h16.npz", context=ctx.numpy(), masks=masks.numpy(),
        text=patch_text.numpy().astype(np.float16),
        token_ids=pre.normalization_stats["token_ids"].numpy().astype(np.int16), scale=pre.normalization_stats["scale"].numpy(),
        encoder_checksum=enc.double().sum(-1).numpy(), generated=tokens.numpy().astype(np.int16), forecast=full.numpy(),
    )
    print("chronos_t5_model", full.shape, tokens[0, :8].tolist())


if __name__ == "__main__":
    # text embeddings are stored as fp16 to keep the fixtures small; the tests up-cast the stored values, so the
    # inputs are identical on both sides.
    timesfm_case("timesfm_l2_b4_c512_h128", 2, 4, 512, 128, padded=True)
    timesfm_case("timesfm_l20_b2_c512_h128", 20, 2, 512, 128, padded=False, seed=1)
    timesfm_case("timesfm_l2_b3_c2048_h64_f2", 2, 3, 2048, 64, padded=True, fusion_layers=2, hidden=(512,), seed=2)
    t5_case()
    chronos2_case()
    chronos_t5_model_case()

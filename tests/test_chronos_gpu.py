"""Chronos-2 adapter path (reference tsfmx/tsfm/chronos.py) on the CUDA path against the restated CPU oracle."""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import chronos2_oracle as C  # noqa: E402  (checker only)
from oracle import timesfm_oracle as O  # noqa: E402
from tsfmx_b200 import ops  # noqa: E402
from tsfmx_b200._lib import DT_BF16, DT_BF16_SPLIT, DT_F32  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module, init_random_  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-3
BF16_TOL = 5e-2


def rel_max(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def build(num_layers, seed=0):
    module = Chronos2Module(num_layers)
    init_random_(module, seed)
    adapter = Chronos2Adapter(module)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    o_model = C.Chronos2Model(C.Chronos2Config(num_layers=num_layers))
    o_adapter = C.OracleChronos2Adapter(o_model)
    o_adapter.load_upstream_state_dict(module.state_dict())
    oracle = O.OracleDecoder(o_adapter, 384, 1, [])
    with torch.no_grad():
        oracle.fusion.projection[0].weight.copy_(dec.fusion.linears()[0].weight)
    return dec.to(DEV).eval(), oracle.eval()


def batch(b, context, horizon, padded, seed=3):
    ctx, masks, _t, _ = O.synthetic_batch(b, context, horizon, padded=padded, seed=seed, patch_len=16)
    g = torch.Generator().manual_seed(seed)
    n = (context + 15) // 16
    text = torch.randn(b, n, 384, generator=g)
    text = text / text.norm(dim=-1, keepdim=True)
    return ctx * 3 + 1, masks, text


def test_encoder_attention_kernel():
    b, t, h, hd = 3, 97, 12, 64
    gen = torch.Generator(device=DEV).manual_seed(0)
    qkv = torch.randn(b * t, 3 * h * hd, generator=gen, device=DEV)
    km = torch.ones(b, t, dtype=torch.bool, device=DEV)
    km[1, :20] = False
    km[2, :] = False  # all keys masked -> uniform
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).to(DEV)
    q, k, v = qkv.double().reshape(b, t, 3, h, hd).permute(2, 0, 3, 1, 4)
    freqs = torch.arange(t, device=DEV).float()[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos().double(), emb.sin().double()
    q = q * cos + C.rotate_half(q) * sin
    k = k * cos + C.rotate_half(k) * sin
    s = q @ k.transpose(-1, -2) + ((~km)[:, None, None, :] * torch.finfo(torch.float32).min).double()
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, h * hd).float()
    got = ops.encoder_attention(qkv, b, t, h, hd, km, inv_freq, DT_F32)
    assert rel_max(got, ref) < 2e-5
    got = ops.encoder_attention(qkv.to(torch.bfloat16), b, t, h, hd, km, inv_freq, DT_BF16_SPLIT)
    assert rel_max(ops.split_to_float(got), ref) < 2e-2


@pytest.mark.parametrize("t", [97, 193, 33, 129, 208, 100])
def test_encoder_attention_tensor_core_path(t):
    """bf16 in / bf16 out runs the mma.sync kernel (one CTA per (series, head)); check it against fp64 torch on the
    same bf16-rounded inputs, including a partly masked and a fully masked series."""
    b, h, hd = 5, 12, 64
    gen = torch.Generator(device=DEV).manual_seed(t)
    qkv = torch.randn(b * t, 3 * h * hd, generator=gen, device=DEV)
    qkv[:, : 2 * h * hd] *= 0.35  # there is no 1/sqrt(d) in this attention: keep the logits O(1) like the model's
    qkv = qkv.to(torch.bfloat16)
    km = torch.ones(b, t, dtype=torch.bool, device=DEV)
    km[1, : t // 3] = False
    km[2, :] = False  # all keys masked -> uniform over all t keys
    km[3, 5::2] = False
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).to(DEV)
    q, k, v = qkv.double().reshape(b, t, 3, h, hd).permute(2, 0, 3, 1, 4)
    freqs = torch.arange(t, device=DEV).float()[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos().double(), emb.sin().double()
    q = q * cos + C.rotate_half(q) * sin
    k = k * cos + C.rotate_half(k) * sin
    s = q @ k.transpose(-1, -2) + ((~km)[:, None, None, :] * torch.finfo(torch.float32).min).double()
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, h * hd).float()
    launches = ops._lib.launch_count()
    got = ops.encoder_attention(qkv, b, t, h, hd, km, inv_freq, DT_BF16)
    assert ops._lib.launch_count() - launches <= 2  # rope table (first call) + the attention kernel
    ops._force_simt_encoder_attention = True
    try:
        simt = ops.encoder_attention(qkv, b, t, h, hd, km, inv_freq, DT_BF16)
    finally:
        ops._force_simt_encoder_attention = False
    assert got.dtype == torch.bfloat16 and got.shape == (b * t, h * hd)
    err, err_simt = rel_max(got.float(), ref), rel_max(got.float(), simt.float())
    assert err < 1e-2, (err, err_simt)  # bf16 q/k after RoPE, bf16 probabilities, bf16 output
    assert err_simt < 1e-2, (err, err_simt)


@pytest.mark.parametrize("padded", [False, True])
def test_preprocess_stage(padded):
    dec, oracle = build(1)
    dec.set_precision("bf16x3")
    ctx, masks, _ = batch(6, 512, 64, padded)
    with torch.no_grad():
        ref = oracle.adapter.preprocess(ctx, masks)
        got = dec.adapter.preprocess(ctx.to(DEV), masks.to(DEV))
    assert torch.equal(got.masks.cpu(), ref.masks)  # bit-exact patch mask
    assert got.normalization_stats["loc"].shape == (6, 1)
    assert (got.normalization_stats["loc"].cpu() - ref.normalization_stats["loc"]).abs().max().item() < 1e-5
    assert rel_max(got.normalization_stats["scale"].cpu(), ref.normalization_stats["scale"]) < 1e-5
    assert rel_max(got.input_embeddings.cpu(), ref.input_embeddings) < 1e-4


@pytest.mark.parametrize("layers,context,horizon,padded", [(2, 512, 128, False), (12, 512, 128, True), (2, 2048, 256, True), (2, 500, 17, False)])
def test_forward_full_parity_fp32_mode(layers, context, horizon, padded):
    dec, oracle = build(layers)
    dec.set_precision("bf16x3")
    ctx, masks, text = batch(5, context, horizon, padded)
    with torch.no_grad():
        ref = oracle.forward_full(horizon, ctx, masks, text)
        got = dec.forward_full(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
        ref_pt = oracle(horizon, ctx, masks, text)
        got_pt = dec(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
    assert got.shape == ref.shape == (5, horizon, 21)
    assert rel_max(got, ref) < FP32_TOL, rel_max(got, ref)
    assert rel_max(got_pt, ref_pt) < FP32_TOL
    assert dec.adapter.point_forecast_index == 10


def test_forward_full_bf16_mode_and_no_text():
    dec, oracle = build(12)
    ctx, masks, text = batch(5, 512, 128, False)
    with torch.no_grad():
        ref = oracle.forward_full(128, ctx, masks, text)
        dec.set_precision("bf16")
        got = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), text.to(DEV)).cpu()
        err = rel_max(got, ref)
        # tolerance derived from the oracle itself evaluated with bf16 weights / activations on the same inputs
        cal = rel_max(O.bf16_oracle(oracle, O.CHRONOS2_BF16_OUTPUTS).forward_full(128, ctx, masks, text), ref)
        print(f"chronos-2 bf16 mode: product rel_max = {err:.3e}, bf16 oracle {cal:.3e}, ratio {err / cal:.2f}")
        assert err < O.BF16_TOL_FACTOR * cal, (err, cal)
        assert err < BF16_TOL
        dec.set_precision("bf16x3")
        ref2 = oracle.forward_full(128, ctx, masks, None)
        got2 = dec.forward_full(128, ctx.to(DEV), masks.to(DEV), None).cpu()
    assert rel_max(got2, ref2) < FP32_TOL


def test_horizon_limit_error():
    dec, oracle = build(1)
    ctx, masks, _ = batch(2, 512, 16, False)
    with pytest.raises(ValueError, match="exceeds the maximum prediction length"):
        dec.forward_full(1025, ctx.to(DEV), masks.to(DEV), None)
    with pytest.raises(ValueError, match="exceeds the maximum prediction length"):
        oracle.forward_full(1025, ctx, masks, None)


# ----------------------------------------------------------------------------- fusion fine-tune through Chronos-2
@pytest.mark.parametrize("t", [97, 40, 130])
def test_encoder_attention_bwd(t):
    """fp32 SIMT backward of the encoder attention core against fp64 torch autograd (partly / fully masked series)."""
    b, h, hd = 4, 12, 64
    gen = torch.Generator(device=DEV).manual_seed(t)
    qkv = torch.randn(b * t, 3 * h * hd, generator=gen, device=DEV)
    qkv[:, : 2 * h * hd] *= 0.35
    dout = torch.randn(b * t, h * hd, generator=gen, device=DEV)
    km = torch.ones(b, t, dtype=torch.bool, device=DEV)
    km[1, : t // 3] = False
    km[2, :] = False
    km[3, 2::3] = False
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).to(DEV)
    qd = qkv.double().requires_grad_(True)
    q, k, v = qd.reshape(b, t, 3, h, hd).permute(2, 0, 3, 1, 4)
    freqs = torch.arange(t, device=DEV).float()[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos().double(), emb.sin().double()
    q = q * cos + C.rotate_half(q) * sin
    k = k * cos + C.rotate_half(k) * sin
    s = q @ k.transpose(-1, -2) + ((~km)[:, None, None, :] * torch.finfo(torch.float32).min).double()
    out = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, h * hd)
    (ref,) = torch.autograd.grad(out, qd, dout.double())
    got = ops.encoder_attention_bwd(qkv, dout, b, t, h, hd, km, inv_freq, DT_F32)
    for name, sl in (("dq", slice(0, h * hd)), ("dk", slice(h * hd, 2 * h * hd)), ("dv", slice(2 * h * hd, None))):
        err = ((got[:, sl].double() - ref[:, sl]).norm() / ref[:, sl].norm().clamp_min(1e-30)).item()
        assert err < 2e-5, (name, err)
    got_s = ops.encoder_attention_bwd(qkv.to(torch.bfloat16), dout.to(torch.bfloat16), b, t, h, hd, km, inv_freq, DT_BF16_SPLIT)
    assert rel_max(ops.split_to_float(got_s), ref.float()) < 3e-2
    # bf16 in / bf16 out: the tensor-core kernel for T <= 112, the SIMT one above
    qb, db = qkv.to(torch.bfloat16), dout.to(torch.bfloat16)
    qd2 = qb.double().requires_grad_(True)
    q2, k2, v2 = qd2.reshape(b, t, 3, h, hd).permute(2, 0, 3, 1, 4)
    q2 = q2 * cos + C.rotate_half(q2) * sin
    k2 = k2 * cos + C.rotate_half(k2) * sin
    s2 = q2 @ k2.transpose(-1, -2) + ((~km)[:, None, None, :] * torch.finfo(torch.float32).min).double()
    out2 = (torch.softmax(s2, -1) @ v2).transpose(1, 2).reshape(b * t, h * hd)
    (ref2,) = torch.autograd.grad(out2, qd2, db.double())
    got_b = ops.encoder_attention_bwd(qb, db, b, t, h, hd, km, inv_freq, DT_BF16)
    assert got_b.dtype == torch.bfloat16
    for name, sl in (("dq", slice(0, h * hd)), ("dk", slice(h * hd, 2 * h * hd)), ("dv", slice(2 * h * hd, None))):
        err = ((got_b[:, sl].double() - ref2[:, sl]).norm() / ref2[:, sl].norm().clamp_min(1e-30)).item()
        assert err < 2e-2, (name, err, t)


@pytest.mark.parametrize("layers,context,horizon,padded", [(2, 512, 64, False), (3, 160, 40, True)])
def test_fusion_gradient_through_chronos2_matches_oracle(layers, context, horizon, padded):
    """MSE of the point forecast -> gradient of the fusion weight, CUDA path (hand-written backward through the frozen
    Chronos-2 encoder) against the restated oracle under torch autograd."""
    dec, oracle = build(layers)
    dec.set_precision("bf16x3")
    dec.adapter.freeze_parameters()
    dec.train()
    ctx, masks, text = batch(6, context, horizon, padded)
    target = torch.randn(6, horizon, generator=torch.Generator().manual_seed(3))
    for p in oracle.fusion.parameters():
        p.requires_grad_(True)
    ref_loss = torch.nn.functional.mse_loss(oracle(horizon, ctx, masks, text), target)
    (ref_grad,) = torch.autograd.grad(ref_loss, [oracle.fusion.projection[0].weight])
    loss = torch.nn.functional.mse_loss(dec(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)), target.to(DEV))
    loss.backward()
    got = dec.fusion.linears()[0].weight.grad.cpu()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    rel = ((got.double() - ref_grad.double()).norm() / ref_grad.double().norm()).item()
    # forecasts carry the 1e-3 bar; for this GRADIENT the measured error is 6e-5 (ctx 512) and 9.7e-4 on the small
    # left-padded case (few valid patches, ReLU gates of near-zero activations decide single terms), so the bound
    # leaves room for one gate falling the other way
    assert rel < 2e-3, rel
    dec.set_precision("bf16")
    dec.zero_grad()
    torch.nn.functional.mse_loss(dec(horizon, ctx.to(DEV), masks.to(DEV), text.to(DEV)), target.to(DEV)).backward()
    got16 = dec.fusion.linears()[0].weight.grad.cpu()
    rel16 = ((got16.double() - ref_grad.double()).norm() / ref_grad.double().norm()).item()
    assert rel16 < 1.5e-1, rel16  # bf16 operands through 2-3 blocks of forward and backward GEMMs; stated separately


def _oracle_grads_by_product_name(model):
    """Gradients of the oracle Chronos-2 model under the upstream (product) state-dict names."""
    g = {"shared.weight": model.shared.weight.grad, "encoder.final_layer_norm.weight": model.final_layer_norm.weight.grad}
    for blk_name, blk in (("input_patch_embedding", model.input_patch_embedding),
                          ("output_patch_embedding", model.output_patch_embedding)):
        for lin in ("hidden_layer", "output_layer", "residual_layer"):
            g[f"{blk_name}.{lin}.weight"] = getattr(blk, lin).weight.grad
            g[f"{blk_name}.{lin}.bias"] = getattr(blk, lin).bias.grad
    for i, blk in enumerate(model.blocks):
        pre = f"encoder.block.{i}.layer."
        for j, (attn, ln) in enumerate(((blk.time_attn, blk.time_ln), (blk.group_attn, blk.group_ln))):
            for proj in "qkvo":
                g[f"{pre}{j}.self_attention.{proj}.weight"] = getattr(attn, proj).weight.grad
            g[f"{pre}{j}.layer_norm.weight"] = ln.weight.grad
        g[pre + "2.mlp.wi.weight"] = blk.wi.weight.grad
        g[pre + "2.mlp.wo.weight"] = blk.wo.weight.grad
        g[pre + "2.layer_norm.weight"] = blk.ff_ln.weight.grad
    return g


@pytest.mark.parametrize("with_text,context,horizon,padded", [(False, 512, 64, False), (True, 160, 40, True)])
def test_full_finetune_gradients_through_chronos2_match_oracle(with_text, context, horizon, padded):
    """The reference's baseline sweep trains either adapter (scripts/tune_baseline_sweep.py:82-87, trainer.py:78-79,123).
    Every Chronos-2 parameter gradient of the CUDA path - weight-gradient GEMMs, norm-scale reductions, the [REG]
    embedding row, the batch-shared future-patch embeddings' share of the input block, the folded W_o W_v of the group
    attention unfolded into dW_o / dW_v (its q / k get exactly zero) - against the restated oracle's torch autograd."""
    dec, oracle = build(2)
    dec.set_precision("bf16x3")
    dec.adapter.unfreeze_parameters()
    dec.train()
    ctx, masks, text = batch(6, context, horizon, padded)
    text_arg = text if with_text else None
    target = torch.randn(6, horizon, generator=torch.Generator().manual_seed(3))
    for p in oracle.parameters():
        p.requires_grad_(True)
    ref_loss = torch.nn.functional.mse_loss(oracle(horizon, ctx, masks, text_arg), target)
    ref_loss.backward()
    ref = _oracle_grads_by_product_name(oracle.adapter._model)
    loss = torch.nn.functional.mse_loss(
        dec(horizon, ctx.to(DEV), masks.to(DEV), None if text_arg is None else text_arg.to(DEV)), target.to(DEV))
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    got = {k: v.grad for k, v in dec.adapter._model.named_parameters()}
    assert set(got) == set(ref)
    worst = {}
    for name, r in ref.items():
        assert got[name] is not None, name
        if r is None or float(r.abs().max()) == 0.0:  # group-attention q / k, the PAD embedding row
            assert float(got[name].abs().max()) == 0.0, name
            continue
        worst[name] = ((got[name].cpu().double() - r.double()).norm() / r.double().norm().clamp_min(1e-30)).item()
    bad = {k: v for k, v in worst.items() if v > 3e-3}
    print("worst chronos-2 full fine-tune gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    assert not bad, bad
    if with_text:
        rf = oracle.fusion.projection[0].weight.grad
        gf = dec.fusion.linears()[0].weight.grad.cpu()
        assert ((gf.double() - rf.double()).norm() / rf.double().norm()).item() < 3e-3


def test_chronos2_forecast_replays_from_a_cuda_graph():
    dec, _ = build(2)
    dec.set_precision("bf16")
    ctx, masks, text = batch(40, 512, 128, True)
    ctx, masks, text = ctx.to(DEV), masks.to(DEV), text.to(DEV)
    with torch.no_grad():
        eager = dec.forward_full(128, ctx, masks, text).clone()
        dec.graphs = True
        try:
            first = dec.forward_full(128, ctx, masks, text).clone()
            ctx.mul_(1.5)
            replay = dec.forward_full(128, ctx, masks, text).clone()
            assert len(dec._graph_cache) == 1
        finally:
            dec.graphs = False
        assert torch.equal(first, eager)
        assert torch.equal(replay, dec.forward_full(128, ctx, masks, text)) and not torch.equal(replay, eager)

"""Fusion fine-tune step (reference tsfmx/trainer.py:200-219) on the CUDA path: backward kernels against torch
autograd, and the fusion-weight gradient / loss curve against the CPU oracle's autograd."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import timesfm_oracle as O  # noqa: E402  (checker only)
from tsfmx_b200 import _lib, ops  # noqa: E402
from tsfmx_b200._lib import ACT_RELU_GRAD, ACT_SILU, ACT_SILU_GRAD, DT_BF16, DT_BF16_SPLIT, DT_F32, PREC_BF16X3  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _rms(v, w, eps=1e-6):
    return w * (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps))


# ----------------------------------------------------------------------------- kernels vs torch autograd (fp64)
@pytest.mark.parametrize("with_v1,with_res,with_v2", [(True, True, True), (False, True, True), (True, True, False), (True, False, True)])
def test_rmsnorm_bwd_chain(with_v1, with_res, with_v2):
    rows, cols = 301, 1280
    gen = torch.Generator(device=DEV).manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=gen, device=DEV)  # noqa: E731
    g_res, v1, g1, v2 = r(rows, cols), r(rows, cols), r(rows, cols), r(rows, cols)
    w1, w2 = 1 + 0.1 * r(cols), 1 + 0.1 * r(cols)
    v1d = v1.double().requires_grad_(True)
    v2d = v2.double().requires_grad_(True)
    total = torch.zeros(rows, cols, dtype=torch.float64, device=DEV)
    if with_v1:
        (gv1,) = torch.autograd.grad(_rms(v1d, w1.double()), v1d, g1.double())
        total = total + gv1
    if with_res:
        total = total + g_res.double()
    g_total = torch.empty(rows, cols, device=DEV)
    g2 = ops.alloc(rows, cols, DT_BF16_SPLIT, torch.device(DEV))
    ops.rmsnorm_bwd_chain(g_res if with_res else None, v1 if with_v1 else None, w1 if with_v1 else None,
                          g1 if with_v1 else None, v2 if with_v2 else None, w2 if with_v2 else None, 1e-6, g_total,
                          DT_BF16_SPLIT, g2 if with_v2 else None, rows, cols)
    assert rel_l2(g_total, total.float()) < 1e-5
    if with_v2:
        (gv2,) = torch.autograd.grad(_rms(v2d, w2.double()), v2d, total)
        assert rel_l2(ops.split_to_float(g2), gv2.float()) < 3e-5


def _attn_fwd_torch(qkv, b, n, h, hd, pm, inv_freq, qw, kw, per_dim):
    d = h * hd
    q, k, v = qkv.reshape(b, n, 3, h, hd).unbind(2)
    nm = pm.sum(-1)
    pos = (torch.arange(n, device=qkv.device)[None, :] - nm[:, None]).double()
    freqs = pos[..., None] * inv_freq.double()[None, None, :]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos()[:, :, None, :], emb.sin()[:, :, None, :]
    rot = lambda t: torch.cat([-t[..., hd // 2 :], t[..., : hd // 2]], -1)  # noqa: E731
    q = q * cos + rot(q) * sin
    k = k * cos + rot(k) * sin
    q = _rms(q, qw.double()) * (torch.nn.functional.softplus(per_dim.double()) * (1.442695041 / math.sqrt(hd)))
    k = _rms(k, kw.double())
    s = torch.einsum("bqhd,bkhd->bhqk", q, k)
    causal = torch.tril(torch.ones(n, n, dtype=torch.bool, device=qkv.device))
    allowed = causal[None, None] & (~pm)[:, None, None, :]
    s = s + torch.where(allowed, 0.0, torch.finfo(torch.float32).min).double()
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(b * n, d)


@pytest.mark.parametrize("n", [16, 5, 40])
def test_attention_bwd(n):
    b, h, hd = 5, 16, 80
    gen = torch.Generator(device=DEV).manual_seed(n)
    qkv = torch.randn(b * n, 3 * h * hd, generator=gen, device=DEV)
    dout = torch.randn(b * n, h * hd, generator=gen, device=DEV)
    pm = torch.zeros(b, n, dtype=torch.bool, device=DEV)
    pm[1, : n // 2] = True
    pm[2, :1] = True
    nm = pm.sum(-1).int()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)).to(DEV)
    qw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    kw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    per_dim = 0.5 * torch.randn(hd, generator=gen, device=DEV)
    q_scale = (torch.nn.functional.softplus(per_dim) * (1.442695041 / math.sqrt(hd))).contiguous()
    qd = qkv.double().requires_grad_(True)
    out = _attn_fwd_torch(qd, b, n, h, hd, pm, inv_freq, qw, kw, per_dim)
    (ref,) = torch.autograd.grad(out, qd, dout.double())
    # all-masked (uniform) rows pass gradient through the additive mask exactly like torch autograd does
    got = ops.timesfm_attention_bwd(qkv, dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_F32)
    assert rel_l2(got, ref.float()) < 2e-5
    got_s = ops.timesfm_attention_bwd(qkv, dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16_SPLIT)
    assert rel_l2(ops.split_to_float(got_s), ref.float()) < 5e-5


@pytest.mark.parametrize("n", [16, 5, 40, 64, 32])
def test_attention_bwd_tensor_core_path(n):
    """bf16 qkv / dO / dqkv run the mma.sync backward kernel (one warp per (series, head), P and dS transposed in
    registers); check against fp64 torch autograd on the same bf16-rounded inputs and against the fp32 SIMT kernel."""
    b, h, hd = 6, 16, 80
    gen = torch.Generator(device=DEV).manual_seed(100 + n)
    qkv = torch.randn(b * n, 3 * h * hd, generator=gen, device=DEV).to(torch.bfloat16)
    dout = torch.randn(b * n, h * hd, generator=gen, device=DEV).to(torch.bfloat16)
    pm = torch.zeros(b, n, dtype=torch.bool, device=DEV)
    pm[1, : n // 2] = True
    pm[2, :1] = True
    pm[3, : n - 1] = True
    nm = pm.sum(-1).int()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)).to(DEV)
    qw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    kw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    per_dim = 0.5 * torch.randn(hd, generator=gen, device=DEV)
    q_scale = (torch.nn.functional.softplus(per_dim) * (1.442695041 / math.sqrt(hd))).contiguous()
    qd = qkv.double().requires_grad_(True)
    out = _attn_fwd_torch(qd, b, n, h, hd, pm, inv_freq, qw, kw, per_dim)
    (ref,) = torch.autograd.grad(out, qd, dout.double())
    got = ops.timesfm_attention_bwd(qkv, dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16)
    _lib.check(_lib.load().tsfmx_attention_force_simt(1))
    try:
        simt = ops.timesfm_attention_bwd(qkv, dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16)
    finally:
        _lib.check(_lib.load().tsfmx_attention_force_simt(0))
    assert got.dtype == torch.bfloat16 and got.shape == (b * n, 3 * h * hd)
    e_ref, e_simt = rel_l2(got.float(), ref.float()), rel_l2(got.float(), simt.float())
    for part, name in ((slice(0, h * hd), "dq"), (slice(h * hd, 2 * h * hd), "dk"), (slice(2 * h * hd, None), "dv")):
        assert rel_l2(got.float()[:, part], ref.float()[:, part]) < 2e-2, (name, e_ref, e_simt)
    assert e_simt < 2e-2, (e_ref, e_simt)


@pytest.mark.parametrize("n,use_bf16", [(16, False), (5, False), (40, False), (16, True), (64, True)])
def test_attention_bwd_parameter_gradients(n, use_bf16):
    """Full fine-tuning needs the gradients of q_ln / k_ln scales and the per-dim scale: both backward kernels return
    d/d(q_ln_w * q_scale) and d/d(k_ln_w); check them (chain rule applied here) against fp64 torch autograd."""
    b, h, hd = 5, 16, 80
    gen = torch.Generator(device=DEV).manual_seed(200 + n)
    qkv = torch.randn(b * n, 3 * h * hd, generator=gen, device=DEV)
    dout = torch.randn(b * n, h * hd, generator=gen, device=DEV)
    if use_bf16:
        qkv, dout = qkv.to(torch.bfloat16), dout.to(torch.bfloat16)
    pm = torch.zeros(b, n, dtype=torch.bool, device=DEV)
    pm[1, : n // 2] = True
    pm[2, :1] = True
    nm = pm.sum(-1).int()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)).to(DEV)
    qw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    kw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    per_dim = 0.5 * torch.randn(hd, generator=gen, device=DEV)
    c = 1.442695041 / math.sqrt(hd)
    q_scale = (torch.nn.functional.softplus(per_dim) * c).contiguous()
    qwd, kwd, pdd = (t.double().requires_grad_(True) for t in (qw, kw, per_dim))
    out = _attn_fwd_torch(qkv.double(), b, n, h, hd, pm, inv_freq, qwd, kwd, pdd)
    g_qw, g_kw, g_pd = torch.autograd.grad(out, [qwd, kwd, pdd], dout.double())
    dparams = torch.zeros(2 * hd, dtype=torch.float32, device=DEV)
    ops.timesfm_attention_bwd(qkv, dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6,
                              DT_BF16 if use_bf16 else DT_F32, dparams=dparams)
    d_eff, d_kw = dparams[:hd].double(), dparams[hd:].double()
    got_qw = d_eff * q_scale.double()
    got_pd = d_eff * qw.double() * c * torch.sigmoid(per_dim.double())
    tol = 3e-2 if use_bf16 else 1e-4
    for name, got, ref in (("q_ln", got_qw, g_qw), ("k_ln", d_kw, g_kw), ("per_dim", got_pd, g_pd)):
        err = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
        assert err < tol, (name, err)


@pytest.mark.parametrize("cols", [1280, 768])
def test_colsum_wgrad(cols):
    rows = 1003
    gen = torch.Generator(device=DEV).manual_seed(cols)
    v = torch.randn(rows, cols, generator=gen, device=DEV) * 2
    g = torch.randn(rows, cols, generator=gen, device=DEV)
    vhat = v.double() * torch.rsqrt(v.double().pow(2).mean(-1, keepdim=True) + 1e-6)
    ref_scale = (g.double() * vhat).sum(0)
    ref_bias = g.double().sum(0)
    got = ops.colsum_wgrad(g, v, 1e-6)
    assert rel_l2(got, ref_scale.float()) < 1e-5
    assert rel_l2(ops.colsum_wgrad(g), ref_bias.float()) < 1e-5
    got16 = ops.colsum_wgrad(g.to(torch.bfloat16), v, 1e-6)
    assert rel_l2(got16, ref_scale.float()) < 1e-2
    acc = ops.colsum_wgrad(g, v, 1e-6, out=got.clone())   # accumulates into `out`
    assert rel_l2(acc, 2 * ref_scale.float()) < 1e-5


def test_gemm_grad_epilogues_and_pre_act():
    m, n, k = 200, 1280, 1280
    gen = torch.Generator(device=DEV).manual_seed(2)
    a32 = torch.randn(m, k, generator=gen, device=DEV)
    b32 = torch.randn(n, k, generator=gen, device=DEV) / math.sqrt(k)
    a, b = ops.cast_rows(a32, DT_BF16_SPLIT), ops.cast_rows(b32, DT_BF16_SPLIT)
    u = torch.randn(m, n, generator=gen, device=DEV)
    acc = a32.double() @ b32.double().t()
    out = torch.empty(m, n, device=DEV)
    ops.gemm([(a, b, k)], m, n, out, DT_F32, precision=PREC_BF16X3, act=ACT_SILU_GRAD, aux=u)
    ud = u.double()
    sg = torch.sigmoid(ud)
    assert rel_l2(out, (acc * sg * (1 + ud * (1 - sg))).float()) < 3e-5
    ops.gemm([(a, b, k)], m, n, out, DT_F32, precision=PREC_BF16X3, act=ACT_RELU_GRAD, aux=u.to(torch.bfloat16))
    assert rel_l2(out, (acc * (u.to(torch.bfloat16).double() > 0)).float()) < 3e-5
    pre = torch.empty(m, n, device=DEV)
    ops.gemm([(a, b, k)], m, n, out, DT_F32, precision=PREC_BF16X3, act=ACT_SILU, pre_act=pre)
    assert rel_l2(pre, acc.float()) < 3e-5
    assert rel_l2(out, (acc * torch.sigmoid(acc)).float()) < 3e-5


@pytest.mark.parametrize("in_kind", ["f32", "bf16", "split"])
def test_transpose_mask_and_mask_cast(in_kind):
    rows, cols = 333, 384
    gen = torch.Generator(device=DEV).manual_seed(3)
    x32 = torch.randn(rows, cols, generator=gen, device=DEV)
    mask = torch.randn(rows, cols, generator=gen, device=DEV)
    if in_kind == "f32":
        x, xf = x32, x32
    elif in_kind == "bf16":
        x = x32.to(torch.bfloat16)
        xf = x.float()
    else:
        x = ops.cast_rows(x32, DT_BF16_SPLIT)
        xf = ops.split_to_float(x)
    out, kpad = ops.transpose_mask(x, rows, cols, DT_BF16_SPLIT, mask=mask)
    assert kpad == 384 and out.shape == (cols, 2 * kpad)
    got = ops.split_to_float(out)
    ref = (xf * (mask > 0)).t()
    assert rel_l2(got[:, :rows], ref) < 1e-5
    assert got[:, rows:].abs().max().item() == 0.0
    out_b, _ = ops.transpose_mask(x, rows, cols, DT_BF16)
    assert torch.equal(out_b[:, :rows], xf.t().to(torch.bfloat16))
    if in_kind == "f32":
        mc = ops.mask_cast_rows(x32, mask, DT_BF16_SPLIT)
        assert rel_l2(ops.split_to_float(mc), x32 * (mask > 0)) < 1e-5


# ----------------------------------------------------------------------------- end to end vs the oracle's autograd
def build(num_layers, fusion_layers=1, hidden=(), seed=0):
    adapter = TimesFM2p5Adapter(num_layers=num_layers, with_quantile_head=False)
    init_random_(adapter, seed=seed)
    torch.manual_seed(seed + 100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, fusion_layers, list(hidden)))
    oracle = O.oracle_from_product(dec)
    return dec.to(DEV), oracle


def oracle_grads(oracle, horizon, ctx, masks, text, target):
    oracle.adapter.freeze_parameters()
    for p in oracle.fusion.parameters():
        p.requires_grad_(True)
        p.grad = None
    loss = torch.nn.functional.mse_loss(oracle(horizon, ctx, masks, text), target)
    loss.backward()
    return loss.item(), [p.grad.clone() for p in oracle.fusion.parameters()]


@pytest.mark.parametrize("layers,fusion_layers,hidden,padded", [(2, 1, (), False), (3, 1, (), True), (2, 2, (512,), False), (2, 3, (256, 128), False)])
def test_fusion_gradient_matches_oracle(layers, fusion_layers, hidden, padded):
    dec, oracle = build(layers, fusion_layers, hidden)
    dec.set_precision("bf16x3")
    dec.adapter.freeze_parameters()
    dec.train()
    ctx, masks, text, target = O.synthetic_batch(6, 512, 64, padded=padded, seed=21)
    ref_loss, ref_grads = oracle_grads(oracle, 64, ctx, masks, text, target)
    point = dec(64, ctx.to(DEV), masks.to(DEV), text.to(DEV))
    loss = torch.nn.functional.mse_loss(point, target.to(DEV))
    loss.backward()
    assert abs(loss.item() - ref_loss) < 1e-4 * max(1.0, abs(ref_loss))
    for lin, ref in zip(dec.fusion.linears(), ref_grads):
        assert lin.weight.grad is not None and lin.weight.grad.shape == ref.shape
        err = rel_l2(lin.weight.grad.cpu(), ref)
        assert err < 1e-3, err  # relative L2 <= 1e-3 in the fp32-accumulate mode
    assert all(p.grad is None for p in dec.adapter.parameters())


def test_gradient_accumulation_and_bf16_mode():
    dec, oracle = build(2)
    dec.adapter.freeze_parameters()
    ctx, masks, text, target = O.synthetic_batch(8, 512, 128, seed=5)
    _, ref_grads = oracle_grads(oracle, 128, ctx, masks, text, target)
    # two micro-batches of 4 with loss / 2 == one batch of 8 (trainer.py:208-209)
    dec.set_precision("bf16x3")
    for sl in (slice(0, 4), slice(4, 8)):
        point = dec(128, ctx[sl].to(DEV), masks[sl].to(DEV), text[sl].to(DEV))
        (torch.nn.functional.mse_loss(point, target[sl].to(DEV)) / 2).backward()
    assert rel_l2(dec.fusion.linears()[0].weight.grad.cpu(), ref_grads[0]) < 1e-3
    dec.zero_grad()
    dec.set_precision("bf16")
    point = dec(128, ctx.to(DEV), masks.to(DEV), text.to(DEV))
    torch.nn.functional.mse_loss(point, target.to(DEV)).backward()
    err = rel_l2(dec.fusion.linears()[0].weight.grad.cpu(), ref_grads[0])
    print(f"bf16-mode fusion gradient rel L2 = {err:.3e}")
    assert err < 6e-2


def test_full_finetune_of_an_adapter_without_that_path_is_refused_loudly():
    """Only the TimesFM adapter has the wgrad path; an unfrozen Chronos-2 adapter must not silently train nothing."""
    from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module
    from tsfmx_b200.tsfm.chronos import init_random_ as init_chronos_

    module = Chronos2Module(1)
    init_chronos_(module, 0)
    dec = MultimodalDecoder(Chronos2Adapter(module), MultimodalDecoderConfig(384, 1, [])).to(DEV).train()
    dec.adapter.unfreeze_parameters()
    ctx, masks, _t, _ = O.synthetic_batch(2, 512, 16, patch_len=16)
    text = torch.randn(2, 32, 384)
    with pytest.raises(NotImplementedError, match="no full fine-tuning path"):
        dec(16, ctx.to(DEV), masks.to(DEV), text.to(DEV))


# ----------------------------------------------------------------------------- full fine-tuning ("baseline" mode)
def _oracle_grads_upstream_names(o_adapter, num_layers):
    """Gradients of the oracle adapter's parameters under the upstream (product) state-dict names."""
    g = {}
    for src, dst in (("hidden_layer", "input_layer"), ("output_layer", "output_layer"), ("residual_layer", "residual_layer")):
        g[f"tokenizer.{src}.weight"] = getattr(o_adapter.tokenizer, dst).weight.grad
        g[f"tokenizer.{src}.bias"] = getattr(o_adapter.tokenizer, dst).bias.grad
        g[f"output_projection_point.{src}.weight"] = getattr(o_adapter.output_projection_point, dst).weight.grad
    for i, layer in enumerate(o_adapter.stacked_xf):
        pre = f"stacked_xf.{i}."
        g[pre + "pre_attn_ln.scale"] = layer.input_layernorm.weight.grad
        g[pre + "post_attn_ln.scale"] = layer.post_attention_layernorm.weight.grad
        g[pre + "pre_ff_ln.scale"] = layer.pre_feedforward_layernorm.weight.grad
        g[pre + "post_ff_ln.scale"] = layer.post_feedforward_layernorm.weight.grad
        a = layer.self_attn
        g[pre + "attn.qkv_proj.weight"] = torch.cat([a.q_proj.weight.grad, a.k_proj.weight.grad, a.v_proj.weight.grad], 0)
        g[pre + "attn.out.weight"] = a.o_proj.weight.grad
        g[pre + "attn.query_ln.scale"] = a.q_norm.weight.grad
        g[pre + "attn.key_ln.scale"] = a.k_norm.weight.grad
        g[pre + "attn.per_dim_scale.per_dim_scale"] = a.scaling.grad
        g[pre + "ff0.weight"] = layer.mlp.fc1.weight.grad
        g[pre + "ff1.weight"] = layer.mlp.fc2.weight.grad
    return g


@pytest.mark.parametrize("with_text,padded", [(False, False), (True, True)])
def test_full_finetune_gradients_match_oracle(with_text, padded):
    """Reference "baseline" mode (trainer.py:78-79,123): the adapter is unfrozen.  Every parameter gradient of the CUDA
    path (weight-gradient GEMMs with K = tokens, column reductions, attention parameter gradients) against the
    oracle's torch autograd, fp32-accumulate parity mode."""
    dec, oracle = build(2)
    dec.set_precision("bf16x3")
    dec.adapter.unfreeze_parameters()
    dec.train()
    ctx, masks, text, target = O.synthetic_batch(6, 512, 128, padded=padded, seed=31)
    text_arg = text if with_text else None
    for p in oracle.parameters():
        p.requires_grad_(True)
    ref_loss = torch.nn.functional.mse_loss(oracle(128, ctx, masks, text_arg), target)
    ref_loss.backward()
    ref = _oracle_grads_upstream_names(oracle.adapter, 2)
    loss = torch.nn.functional.mse_loss(
        dec(128, ctx.to(DEV), masks.to(DEV), None if text_arg is None else text_arg.to(DEV)), target.to(DEV))
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    got = {k: v.grad for k, v in dec.adapter._model.named_parameters()}
    worst = {}
    for name, r in ref.items():
        assert got[name] is not None, name
        worst[name] = ((got[name].cpu().double() - r.double()).norm() / r.double().norm().clamp_min(1e-30)).item()
    bad = {k: v for k, v in worst.items() if v > 2e-3}
    assert not bad, bad
    if with_text:
        rf = oracle.fusion.projection[0].weight.grad
        gf = dec.fusion.linears()[0].weight.grad.cpu()
        assert ((gf.double() - rf.double()).norm() / rf.double().norm()).item() < 2e-3
    else:
        assert dec.fusion.linears()[0].weight.grad is None

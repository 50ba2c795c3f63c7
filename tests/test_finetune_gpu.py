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


@pytest.mark.parametrize("rows,cols", [(1003, 384), (64, 3072), (5, 3072), (70000, 336), (3, 4)])
def test_colsum_bias_any_width(rows, cols):
    """Bias gradients of the blocks that are not model-wide (Chronos-2: 3072-wide hidden layers, the 336-wide head)."""
    gen = torch.Generator(device=DEV).manual_seed(rows + cols)
    g = torch.randn(rows, cols, generator=gen, device=DEV)
    ref = g.double().sum(0).float()
    assert rel_l2(ops.colsum_wgrad(g), ref) < 1e-5
    assert rel_l2(ops.colsum_wgrad(g.to(torch.bfloat16)), g.to(torch.bfloat16).double().sum(0).float()) < 1e-5


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
    """TimesFM and Chronos-2 have the wgrad path; an unfrozen Chronos-T5 adapter (not a reference adapter) must not
    silently train nothing."""
    from tsfmx_b200.tsfm.chronos_t5 import ChronosT5Adapter, ChronosT5Module
    from tsfmx_b200.tsfm.chronos_t5 import init_random_ as init_t5_

    module = ChronosT5Module(num_layers=1)
    init_t5_(module, 0)
    adapter = ChronosT5Adapter(module)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(DEV).train()
    dec.adapter.unfreeze_parameters()
    ctx, masks, text, _ = O.synthetic_batch(2, 64, 8)
    with pytest.raises(NotImplementedError, match="no full fine-tuning path"):
        dec(8, ctx.to(DEV), masks.to(DEV), adapter.expand_text_embeddings(text, 64).to(DEV))


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


# ------------------------------------------------------------------------------------------------ loss curves
def _samples(n, horizon=64, seed=77):
    ctx, _m, text, hor = O.synthetic_batch(n, 512, horizon, seed=seed)
    return [{"context": ctx[i].numpy(), "horizon": hor[i].numpy(), "text_embeddings": text[i].numpy(), "metadata": {"i": i}}
            for i in range(n)]


def _train_args(tmp_path, lr, **kw):
    import types

    base = dict(per_device_train_batch_size=4, per_device_eval_batch_size=4, gradient_accumulation_steps=1,
                max_grad_norm=1.0, learning_rate=lr, weight_decay=0.01, num_train_epochs=1, seed=3, warmup_steps=0.25,
                lr_scheduler_type="linear", save_strategy="no", checkpoint_dir=tmp_path / "ckpt")
    base.update(kw)
    return types.SimpleNamespace(**base)


def _reference_loop(oracle, params, samples, args, collate, with_text):
    """The reference's train_epoch (trainer.py:186-231) restated on the CPU oracle: same DataLoader order (same seeded
    generator), MSE on the point forecast, clip, AdamW, linear warm-up schedule stepped per optimizer step."""
    from torch.utils.data import DataLoader

    from tsfmx_b200.trainer import linear_schedule_with_warmup

    loader = DataLoader(samples, batch_size=args.per_device_train_batch_size, shuffle=True, num_workers=0, collate_fn=collate,
                        generator=torch.Generator().manual_seed(args.seed))
    total = args.num_train_epochs * len(loader)
    opt = torch.optim.AdamW(params, lr=args.learning_rate, weight_decay=args.weight_decay)
    import math
    sched = linear_schedule_with_warmup(opt, math.ceil(total * args.warmup_steps), total)
    losses = []
    for batch in loader:
        ctx, hor = batch["context"], batch["horizon"]
        pad = torch.zeros_like(ctx, dtype=torch.bool)
        point = oracle(hor.shape[-1], ctx, pad, batch["text_embeddings"] if with_text else None)
        loss = torch.nn.functional.mse_loss(point, hor)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, args.max_grad_norm)
        opt.step()
        opt.zero_grad()
        sched.step()
        losses.append(loss.item())
    return losses


@pytest.mark.parametrize("graphs", [False, True])
def test_fusion_finetune_loss_curve_matches_oracle(tmp_path, graphs):
    """identical loss curve for >= 10 optimizer steps with the same seed and batches.  12 steps
    of the reference's fine-tune loop (frozen adapter, trainable fusion) through MultimodalTrainer - eager, and with the
    forward + backward replayed from a CUDA graph from the second batch on - against oracle autograd + torch AdamW."""
    from tsfmx_b200.data import multimodal_collate_fn
    from tsfmx_b200.trainer import MultimodalTrainer

    dec, oracle = build(2)
    dec.set_precision("bf16x3")
    samples = _samples(48)
    args = _train_args(tmp_path, lr=2e-3, cuda_graphs=graphs)
    oracle.adapter.freeze_parameters()
    ref_params = list(oracle.fusion.parameters())
    for p in ref_params:
        p.requires_grad_(True)
    ref_losses = _reference_loop(oracle, ref_params, samples, args, multimodal_collate_fn, True)
    trainer = MultimodalTrainer(dec, args, samples, samples[:8], "multimodal", torch.device(DEV))
    got = []
    micro = trainer._micro_batch
    trainer._micro_batch = lambda batch, accum: got.append(micro(batch, accum)) or got[-1]
    epoch_loss = trainer.train_epoch()
    got = [float(x) for x in got]
    assert len(got) == len(ref_losses) == 12 and trainer.global_step == 12
    assert trainer.graph_replays == (11 if graphs else 0)
    print("loss curve (product / oracle):", [f"{a:.5f}/{b:.5f}" for a, b in zip(got, ref_losses)])
    for a, b in zip(got, ref_losses):
        assert a == pytest.approx(b, rel=1e-3), (got, ref_losses)
    assert ref_losses[-1] < ref_losses[0]  # and it is actually learning
    assert epoch_loss == pytest.approx(sum(ref_losses) / 12, rel=1e-3)
    w_ref = ref_params[0].detach()
    assert rel_l2(dec.fusion.linears()[0].weight.detach().cpu(), w_ref) < 1e-3


def test_full_finetune_loss_curve_matches_oracle(tmp_path):
    """The same for the reference's "baseline" mode (trainer.py:78-79: every adapter parameter trained, no text): 10
    optimizer steps, graph replay from the second batch on."""
    from tsfmx_b200.data import baseline_collate_fn
    from tsfmx_b200.trainer import MultimodalTrainer

    dec, oracle = build(2)
    dec.set_precision("bf16x3")
    samples = [{k: v for k, v in s.items() if k != "text_embeddings"} for s in _samples(40, seed=78)]
    args = _train_args(tmp_path, lr=2e-4, cuda_graphs=True)
    oracle.adapter.unfreeze_parameters()
    ref_params = list(oracle.adapter.parameters())
    ref_losses = _reference_loop(oracle, ref_params, samples, args, baseline_collate_fn, False)
    trainer = MultimodalTrainer(dec, args, samples, samples[:8], "baseline", torch.device(DEV))
    got = []
    micro = trainer._micro_batch
    trainer._micro_batch = lambda batch, accum: got.append(micro(batch, accum)) or got[-1]
    trainer.train_epoch()
    got = [float(x) for x in got]
    assert len(got) == len(ref_losses) == 10 and trainer.graph_replays == 9
    print("full fine-tune loss curve (product / oracle):", [f"{a:.5f}/{b:.5f}" for a, b in zip(got, ref_losses)])
    for a, b in zip(got, ref_losses):
        assert a == pytest.approx(b, rel=2e-3), (got, ref_losses)
    assert ref_losses[-1] < ref_losses[0]


def test_trainer_full_loop_with_graphs_validation_and_checkpoints(tmp_path):
    """MultimodalTrainer.train() end to end on the GPU: two epochs of graph-replayed training steps interleaved with
    eval-mode validation passes (which re-pack the just-updated fusion weights outside the graph), best-model
    checkpointing and the reload at the end - the interplay of the training graph, the weight caches and the eager
    forecast path."""
    from tsfmx_b200.trainer import MultimodalTrainer

    dec, _ = build(2)
    dec.set_precision("bf16")
    train, val = _samples(32, seed=91), _samples(8, seed=92)
    args = _train_args(tmp_path, lr=2e-3, cuda_graphs=True, num_train_epochs=3, save_strategy="best",
                       eval_strategy="epoch", load_best_model_at_end=True, logging_strategy="epoch")
    logged = []

    class Run:
        def log(self, metrics, step):
            logged.append((step, dict(metrics)))

    trainer = MultimodalTrainer(dec, args, train, val, "multimodal", torch.device(DEV), wandb_run=Run())
    w0 = dec.fusion.linears()[0].weight.detach().clone()
    before = trainer.validate_epoch()
    trainer.train()
    assert trainer.global_step == 3 * 8 and trainer.graph_replays == 3 * 8 - 1
    assert [s for s, _ in logged] == [8, 16, 24]
    vals = [m["val/loss"] for _, m in logged]
    assert min(vals) < before and trainer.best_val_loss == pytest.approx(min(vals))
    assert all(m["train/loss"] == m["train/loss"] for _, m in logged)  # no NaN
    best = torch.load(args.checkpoint_dir / "best_model.pt", weights_only=True)
    assert best["best_val_loss"] == pytest.approx(min(vals))
    # the best fusion weights are back in the model, and an eager validation pass reproduces their loss
    assert torch.equal(dec.fusion.linears()[0].weight.detach().cpu(), best["fusion_state_dict"]["projection.0.weight"].cpu())
    assert not torch.equal(dec.fusion.linears()[0].weight.detach(), w0)
    assert trainer.validate_epoch() == pytest.approx(min(vals), rel=1e-4)
    trainer.release_graphs()
    assert not trainer._train_graphs

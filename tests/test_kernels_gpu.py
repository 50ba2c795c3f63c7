"""Kernel-level GPU tests: each C-ABI entry point against a plain PyTorch fp32 restatement of the same op.

(The model-level parity tests against the oracle live in test_parity_gpu.py.)
"""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from tsfmx_b200 import _lib, ops  # noqa: E402
from tsfmx_b200._lib import (  # noqa: E402
    ACT_NONE,
    ACT_RELU,
    ACT_SILU,
    DT_BF16,
    DT_BF16_SPLIT,
    DT_F32,
    PREC_BF16,
    PREC_BF16X3,
)

DEV = "cuda"


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _to_float(t: torch.Tensor, dtype: int) -> torch.Tensor:
    if dtype == DT_BF16_SPLIT:
        return ops.split_to_float(t)
    return t.float()


# ----------------------------------------------------------------------------- GEMM
def _make_operand(rows, k, precision, gen):
    x = torch.randn(rows, k, generator=gen, device=DEV, dtype=torch.float32)
    if precision == PREC_BF16:
        xb = x.to(torch.bfloat16)
        return xb, xb.float()
    xs = ops.cast_rows(x, DT_BF16_SPLIT)
    return xs, x


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize(
    "m,n,k",
    [(128, 256, 64), (256, 256, 128), (1000, 1280, 384), (4096, 3840, 1280), (300, 336, 768), (64, 1280, 1280)],
)
def test_gemm_bf16_plain(cta_group, m, n, k):
    lib = _lib.load()
    _lib.check(lib.tsfmx_gemm_set_cta_group(cta_group))
    try:
        gen = torch.Generator(device=DEV).manual_seed(m * 7 + n * 3 + k)
        a, af = _make_operand(m, k, PREC_BF16, gen)
        b, bf = _make_operand(n, k, PREC_BF16, gen)
        out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32)
        ops.gemm([(a, b, k)], m, n, out, DT_F32)
        ref = af @ bf.t()
        assert _rel(out, ref) < 2e-5, (cta_group, m, n, k)
    finally:
        _lib.check(lib.tsfmx_gemm_set_cta_group(0))


@pytest.mark.parametrize("n,act,d_dtype", [(3840, ACT_NONE, DT_BF16), (1280, ACT_SILU, DT_BF16), (1280, ACT_NONE, DT_F32)])
def test_gemm_at_the_benchmarked_size(n, act, d_dtype):
    """The decoder-layer GEMMs exactly as bench.py launches them: M = 65 536 token rows (4096 series x 16 patches),
    K = 1280, N = 3840 (qkv) / 1280 (ff0 with the SiLU epilogue, attn-out / ff1), every one of the 256 x 15 (or 5) tiles
    checked against torch's fp32 matmul of the same bf16 operands."""
    torch.backends.cuda.matmul.allow_tf32 = False
    m, k = 65536, 1280
    gen = torch.Generator(device=DEV).manual_seed(n + act)
    a, af = _make_operand(m, k, PREC_BF16, gen)
    b, bf = _make_operand(n, k, PREC_BF16, gen)
    af *= 0.05
    a = af.to(torch.bfloat16)
    af = a.float()
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32 if d_dtype == DT_F32 else torch.bfloat16)
    ops.gemm([(a, b, k)], m, n, out, d_dtype, act=act)
    ref = af @ bf.t()
    if act == ACT_SILU:
        ref = torch.nn.functional.silu(ref)
    tol = 2e-5 if d_dtype == DT_F32 else 5e-3  # bf16 storage of the output: 2^-9 relative per element
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("split", [0, 3, 16])
@pytest.mark.parametrize("m,n,k", [(1280, 1280, 16384), (3840, 1280, 4096), (300, 336, 8192)])
def test_gemm_split_k(cta_group, split, m, n, k):
    """Weight-gradient shapes (few output tiles, K = tokens): split-K tiles add their partial sums with reductions."""
    lib = _lib.load()
    _lib.check(lib.tsfmx_gemm_set_cta_group(cta_group))
    _lib.check(lib.tsfmx_gemm_set_split_k(split))
    try:
        gen = torch.Generator(device=DEV).manual_seed(m + n + k + split)
        a, af = _make_operand(m, k, PREC_BF16, gen)
        b, bf = _make_operand(n, k, PREC_BF16, gen)
        ref = (af.double() @ bf.double().t()).float()
        out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32)
        ops.gemm([(a, b, k)], m, n, out, DT_F32)
        assert _rel(out, ref) < 2e-5
        # in-place accumulation (residual = D): D <- D + A B^T
        acc = torch.randn(m, n, generator=gen, device=DEV)
        want = acc + ref
        ops.gemm([(a, b, k)], m, n, acc, DT_F32, residual=acc)
        assert _rel(acc, want) < 2e-5
        # fp32-accurate mode splits its three segments the same way
        a3, af3 = _make_operand(m, k, PREC_BF16X3, gen)
        b3, bf3 = _make_operand(n, k, PREC_BF16X3, gen)
        out3 = torch.empty(m, n, device=DEV, dtype=torch.float32)
        ops.gemm([(a3, b3, k)], m, n, out3, DT_F32, precision=PREC_BF16X3)
        assert _rel(out3, (af3.double() @ bf3.double().t()).float()) < 3e-5
    finally:
        _lib.check(lib.tsfmx_gemm_set_cta_group(0))
        _lib.check(lib.tsfmx_gemm_set_split_k(0))


@pytest.mark.parametrize("split", [0, 1, 5])
@pytest.mark.parametrize(
    "rows,n_out,k_in",
    [(4096 + 37, 1280, 1280), (16384, 3840, 1280), (1000, 336, 768), (513, 64, 1280), (200, 1280, 64), (70, 8, 264)],
)
def test_wgrad_token_major(split, rows, n_out, k_in):
    """dW = dY^T X with both operands read as they lie in HBM ([tokens, features] = MN-major UMMA descriptors, no
    transposed copies): ragged token counts (TMA zero-fills the last k-block), feature counts off the 64-element chunk
    and the 256-wide tile, strided operands, with and without split-K - against torch fp32 and against the transposing
    path (same products, same accumulation order inside a tile)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.load()
    _lib.check(lib.tsfmx_gemm_set_split_k(split))
    try:
        gen = torch.Generator(device=DEV).manual_seed(rows + n_out)
        dy_full = torch.randn(rows, n_out + 24, device=DEV, generator=gen).to(torch.bfloat16)
        dy = dy_full[:, 8 : 8 + n_out]  # strided rows, 16-byte aligned start
        x = torch.randn(rows, k_in, device=DEV, generator=gen).to(torch.bfloat16)
        got = ops.wgrad(dy, x, rows, n_out, k_in, PREC_BF16)
        ref = dy.float().t() @ x.float()
        assert _rel(got, ref) < 3e-5
        dy_t, kpad = ops.transpose_mask(dy.contiguous(), rows, n_out, DT_BF16)
        x_t, _ = ops.transpose_mask(x, rows, k_in, DT_BF16)
        old = torch.empty(n_out, k_in, dtype=torch.float32, device=DEV)
        ops.gemm([(dy_t, x_t, kpad)], n_out, k_in, old, DT_F32, precision=PREC_BF16)
        if split == 1:  # one tile per output block: the two paths add the same products in the same order
            assert torch.equal(got, old)
        else:
            assert _rel(got, old) < 1e-5
    finally:
        _lib.check(lib.tsfmx_gemm_set_split_k(0))


def test_wgrad_falls_back_for_fp32_and_split_operands():
    """Parity mode (split operands) and fp32 gradients keep the transposing path; both give the fp32 product."""
    gen = torch.Generator(device=DEV).manual_seed(5)
    rows, n_out, k_in = 777, 128, 192
    dy = torch.randn(rows, n_out, device=DEV, generator=gen)
    x = torch.randn(rows, k_in, device=DEV, generator=gen)
    ref = dy.t() @ x
    got = ops.wgrad(ops.cast_rows(dy, DT_BF16_SPLIT), ops.cast_rows(x, DT_BF16_SPLIT), rows, n_out, k_in, PREC_BF16X3)
    assert _rel(got, ref) < 1e-4
    got = ops.wgrad(dy, x.to(torch.bfloat16), rows, n_out, k_in, PREC_BF16)  # fp32 dY: cast while transposing
    assert _rel(got, dy.to(torch.bfloat16).float().t() @ x.to(torch.bfloat16).float()) < 3e-5


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_bf16x3_close_to_fp32(cta_group):
    lib = _lib.load()
    _lib.check(lib.tsfmx_gemm_set_cta_group(cta_group))
    try:
        m, n, k = 512, 1280, 1280
        gen = torch.Generator(device=DEV).manual_seed(5)
        a, af = _make_operand(m, k, PREC_BF16X3, gen)
        b, bf = _make_operand(n, k, PREC_BF16X3, gen)
        out = torch.empty(m, n, device=DEV, dtype=torch.float32)
        ops.gemm([(a, b, k)], m, n, out, DT_F32, precision=PREC_BF16X3)
        ref = (af.double() @ bf.double().t()).float()
        assert _rel(out, ref) < 3e-5
    finally:
        _lib.check(lib.tsfmx_gemm_set_cta_group(0))


@pytest.mark.parametrize("precision", [PREC_BF16, PREC_BF16X3])
@pytest.mark.parametrize("d_dtype", [DT_F32, DT_BF16, DT_BF16_SPLIT])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_SILU, ACT_RELU])
def test_gemm_epilogue(precision, d_dtype, act):
    m, n, k1, k2 = 384, 1280, 1280, 64
    gen = torch.Generator(device=DEV).manual_seed(11 + act + 10 * d_dtype)
    a1, a1f = _make_operand(m, k1, precision, gen)
    b1, b1f = _make_operand(n, k1, precision, gen)
    a2, a2f = _make_operand(m, k2, precision, gen)
    b2, b2f = _make_operand(n, k2, precision, gen)
    bias = torch.randn(n, generator=gen, device=DEV)
    rs = torch.rand(m, generator=gen, device=DEV) + 0.5
    rsh = torch.randn(m, generator=gen, device=DEV)
    res = torch.randn(m, n, generator=gen, device=DEV)
    out = ops.alloc(m, n, d_dtype, torch.device(DEV))
    ops.gemm(
        [(a1, b1, k1), (a2, b2, k2)], m, n, out, d_dtype, precision=precision, act=act, bias=bias, row_scale=rs,
        row_shift=rsh, residual=res,
    )
    acc = (a1f.double() @ b1f.double().t() + a2f.double() @ b2f.double().t()) / math.sqrt(k1)
    # (scale the comparison, not the op: keep magnitudes O(1) for the activation)
    v = a1f.double() @ b1f.double().t() + a2f.double() @ b2f.double().t() + bias.double()
    if act == ACT_SILU:
        v = v * torch.sigmoid(v)
    elif act == ACT_RELU:
        v = v.clamp_min(0)
    v = v * rs.double()[:, None] + rsh.double()[:, None] + res.double()
    got = _to_float(out, d_dtype)
    tol = {DT_F32: 3e-5, DT_BF16: 6e-3, DT_BF16_SPLIT: 5e-5}[d_dtype]
    if precision == PREC_BF16 and d_dtype != DT_BF16:
        tol = 3e-5
    assert _rel(got, v.float()) < tol
    del acc


def test_gemm_n_store_and_ldd():
    # head epilogue: only the first h*q columns are kept, output row stride = h*q (not a multiple of 4)
    m, n, k = 200, 1280, 1280
    gen = torch.Generator(device=DEV).manual_seed(3)
    a, af = _make_operand(m, k, PREC_BF16, gen)
    b, bf = _make_operand(n, k, PREC_BF16, gen)
    for n_store in (70, 1280, 640):
        out = torch.full((m, n_store), float("nan"), device=DEV, dtype=torch.float32)
        ops.gemm([(a, b, k)], m, n, out, DT_F32, n_store=n_store)
        ref = (af @ bf.t())[:, :n_store]
        assert _rel(out, ref) < 2e-5


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(128, 256, device=DEV)
    with pytest.raises(_lib.TsfmxError):
        ops.gemm([(a, a, 48)], 128, 128, out, DT_F32)  # k not a multiple of 64


# ----------------------------------------------------------------------------- TimesFM preprocess
def _ref_patchify(x, mask, p=32):
    b, c = x.shape
    n = c // p
    xp = x.reshape(b, n, p).double()
    mp = mask.reshape(b, n, p)
    cnt = torch.zeros(b, dtype=torch.float64, device=x.device)
    mean = torch.zeros_like(cnt)
    std = torch.zeros_like(cnt)
    mus, sigmas = [], []
    for i in range(n):
        valid = (~mp[:, i]).double()
        ic = valid.sum(-1)
        ics = torch.where(ic == 0, torch.ones_like(ic), ic)
        im = torch.where(ic == 0, torch.zeros_like(ic), (xp[:, i] * valid).sum(-1) / ics)
        iv = torch.where(ic == 0, torch.zeros_like(ic), (((xp[:, i] - im[:, None]) * valid) ** 2).sum(-1) / ics)
        isd = iv.clamp_min(0).sqrt()
        nc = cnt + ic
        ncs = torch.where(nc == 0, torch.ones_like(nc), nc)
        nm = torch.where(nc == 0, torch.zeros_like(nc), (cnt * mean + im * ic) / ncs)
        nv = (cnt * std**2 + ic * isd**2 + cnt * (mean - nm) ** 2 + ic * (im - nm) ** 2) / ncs
        nv = torch.where(nc == 0, torch.zeros_like(nv), nv)
        cnt, mean, std = nc, nm, nv.clamp_min(0).sqrt()
        mus.append(mean)
        sigmas.append(std)
    mu = torch.stack(mus, 1)
    sigma = torch.stack(sigmas, 1)
    safe = torch.where(sigma < 1e-6, torch.ones_like(sigma), sigma)
    normed = (xp - mu[..., None]) / safe[..., None]
    normed = torch.where(mp, torch.zeros_like(normed), normed)
    tokens = torch.cat([normed, mp.double()], -1).reshape(b * n, 2 * p)
    return tokens.float(), mu.float(), sigma.float(), mp[..., -1]


def _random_left_padding(b, c, gen):
    pad = torch.randint(0, c, (b,), generator=gen, device=DEV)
    pad[0] = 0
    pad[1] = 37  # partial patch
    pad[2] = 64  # two full patches
    pad[3] = c  # everything padded
    return torch.arange(c, device=DEV)[None, :] < pad[:, None]


@pytest.fixture
def generic_kernels():
    """Force the generic fallback kernels (what rows longer than 4096 / unaligned rows / odd contexts take) through the
    library's tuning hook, so that they stay covered next to the staged default kernels."""
    lib = ops._lib.load()
    ops._lib.check(lib.tsfmx_tune(2, 1))
    ops._lib.check(lib.tsfmx_tune(3, 1))
    yield
    ops._lib.check(lib.tsfmx_tune(2, 0))
    ops._lib.check(lib.tsfmx_tune(3, 0))


def test_fallback_kernels_agree_with_the_staged_ones(generic_kernels):
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(67, 512, generator=gen, device=DEV) * 3 + 1
    mask = _random_left_padding(67, 512, gen)
    generic = ops.timesfm_patchify_norm(x, mask, 32, DT_F32)
    ops._lib.check(ops._lib.load().tsfmx_tune(3, 0))
    staged = ops.timesfm_patchify_norm(x, mask, 32, DT_F32)
    assert torch.equal(generic[3], staged[3]) and torch.equal(generic[4], staged[4])
    for g, s_ in zip(generic[:3], staged[:3]):
        assert (g - s_).abs().max().item() < 2e-6 * max(1.0, s_.abs().max().item())
    _centers, boundaries = _t5_tables()
    xt = torch.randn(129, 512, generator=torch.Generator().manual_seed(1)).to(DEV) * 4
    generic_ids = ops.chronos_t5_tokenize(xt, boundaries.to(DEV))
    ops._lib.check(ops._lib.load().tsfmx_tune(2, 0))
    staged_ids = ops.chronos_t5_tokenize(xt, boundaries.to(DEV))
    for g, s_ in zip(generic_ids, staged_ids):
        assert torch.equal(g, s_)


@pytest.mark.parametrize("context", [32, 512, 2048, 96, 1056, 4096, 8192])
@pytest.mark.parametrize("tokens_dtype", [DT_F32, DT_BF16, DT_BF16_SPLIT])
def test_timesfm_patchify_norm(context, tokens_dtype):
    b = 67
    gen = torch.Generator(device=DEV).manual_seed(context)
    x = torch.randn(b, context, generator=gen, device=DEV) * 3 + 1
    x[5] = 2.5  # constant series: sigma == 0 -> safe sigma path
    mask = _random_left_padding(b, context, gen)
    mask[6, ::7] = True  # scattered (non left) padding as well
    tokens, mu, sigma, pm, nm = ops.timesfm_patchify_norm(x, mask, 32, tokens_dtype)
    rt, rmu, rsig, rpm = _ref_patchify(x, mask)
    assert torch.equal(pm, rpm)  # bit-exact patch mask
    assert torch.equal(nm.long(), rpm.sum(-1))
    # 2e-6 absolute at the headline contexts; a few more ulp (relative to values of ~3) after 256 merges at ctx 8192
    stat_tol = 2e-6 if context <= 4096 else 2e-6 * float(rsig.abs().max().clamp_min(1.0))
    assert (mu - rmu).abs().max().item() < stat_tol
    assert (sigma - rsig).abs().max().item() < stat_tol
    got = _to_float(tokens, tokens_dtype)
    # mask half of the token is exact in every storage type
    assert torch.equal(got[:, 32:], rt[:, 32:])
    tol = 1e-2 if tokens_dtype == DT_BF16 else 2e-4
    # the sigma < 1e-6 -> 1 switch is discontinuous: compare where the reference sigma is clear of it
    ok = ((rsig - 1e-6).abs() > 1e-7).reshape(-1)
    assert (got[ok, :32] - rt[ok, :32]).abs().max().item() < tol * max(1.0, rt[ok, :32].abs().max().item())


def test_timesfm_patchify_norm_errors():
    x = torch.zeros(4, 500, device=DEV)
    m = torch.zeros(4, 500, dtype=torch.bool, device=DEV)
    with pytest.raises(_lib.TsfmxError, match="divisible"):
        ops.timesfm_patchify_norm(x, m)
    with pytest.raises(_lib.TsfmxError):
        ops.timesfm_patchify_norm(x.cpu(), m.cpu())


# ----------------------------------------------------------------------------- Chronos-T5 tokeniser
def _t5_tables():
    centers = torch.linspace(-15.0, 15.0, 4096 - 2 - 1)
    boundaries = torch.cat([torch.tensor([-1e20]), (centers[1:] + centers[:-1]) / 2, torch.tensor([1e20])])
    return centers, boundaries


@pytest.mark.parametrize("context", [512, 2048, 1024, 516, 1540, 8192, 100, 36, 4099])
def test_chronos_t5_tokenize_bit_exact(context):
    b = 129
    gen = torch.Generator().manual_seed(context)
    x = torch.randn(b, context, generator=gen) * torch.rand(b, 1, generator=gen) * 10
    x[0, :5] = float("nan")
    x[1] = float("nan")  # fully missing series -> scale 1, all PAD
    x[2] = 0.0  # scale not > 0 -> 1
    x[3] *= 0.01
    x[3, 10], x[3, 20] = 1e6, -1e6  # outliers -> clamp to the outer bins
    centers, boundaries = _t5_tables()
    # exact hits on bin boundaries (right=True tie rule)
    x[4, : min(context, 4094)] = boundaries[: min(context, 4094)].clamp(-1e4, 1e4)
    am = ~torch.isnan(x)
    scale = (torch.nansum(x.abs().double() * am, -1).float() / torch.nansum(am.float(), -1))
    scale[~(scale > 0)] = 1.0
    ref = torch.bucketize(x / scale[:, None], boundaries, right=True) + 2
    ref.clamp_(0, 4095)
    ref[~am] = 0
    ref = torch.cat([ref, torch.ones(b, 1, dtype=torch.long)], 1)
    ref_am = torch.cat([am, torch.ones(b, 1, dtype=torch.bool)], 1)
    ids, gam, gscale = ops.chronos_t5_tokenize(x.to(DEV), boundaries.to(DEV))
    assert torch.equal(gscale.cpu(), scale)
    assert torch.equal(gam.cpu(), ref_am)
    assert torch.equal(ids.cpu(), ref)
    vals = ops.chronos_t5_dequantize(ids, centers.to(DEV), gscale)
    ref_vals = centers[(ref - 3).clamp(0, 4092)] * scale[:, None]
    assert torch.equal(vals.cpu(), ref_vals)


def test_chronos_t5_tokenize_division_stress():
    """The kernel divides by the per-series scale with a hoisted reciprocal + two FMA refinements; every id must still
    equal torch's IEEE division + bucketize, including values parked on / next to bin boundaries and series whose
    scale spans 80 binades."""
    b, context = 4096, 512
    gen = torch.Generator().manual_seed(7)
    expo = torch.randint(-40, 40, (b, 1), generator=gen).float()
    x = torch.randn(b, context, generator=gen) * torch.exp2(expo)
    centers, boundaries = _t5_tables()
    # rows 0..1023: every element sits exactly on / one ulp around (boundary * scale) of a first-pass scale
    inner = boundaries[1:-1]
    am = torch.ones_like(x, dtype=torch.bool)
    scale0 = (x.abs().double().sum(-1).float() / context)
    pick = inner[torch.randint(0, inner.numel(), (1024, context), generator=gen)]
    edge = pick * scale0[:1024, None]
    nudged = torch.nextafter(edge, torch.where(torch.rand(1024, context, generator=gen) < 0.5, -1.0, 1.0) * torch.inf)
    x[:1024] = torch.where(torch.rand(1024, context, generator=gen) < 0.5, edge, nudged)
    scale = (x.abs().double().sum(-1).float() / context)
    scale[~(scale > 0)] = 1.0
    ref = torch.bucketize(x / scale[:, None], boundaries, right=True) + 2
    ref.clamp_(0, 4095)
    ids, gam, gscale = ops.chronos_t5_tokenize(x.to(DEV), boundaries.to(DEV))
    assert torch.equal(gscale.cpu(), scale)
    assert torch.equal(ids.cpu()[:, :context], ref)
    assert bool(gam.all())


# ----------------------------------------------------------------------------- Chronos-2 context preparation
@pytest.mark.parametrize("context", [512, 2048, 500])
@pytest.mark.parametrize("out_cols", [48, 64])
def test_chronos2_patchify_norm(context, out_cols):
    b, p = 33, 16
    gen = torch.Generator().manual_seed(context)
    x = torch.randn(b, context, generator=gen) * 4 + 2
    mask = torch.zeros(b, context, dtype=torch.bool)
    mask[1, :100] = True
    mask[2, :] = True
    x[3] = 1.5  # zero variance -> eps scale
    n = (context + p - 1) // p
    pad = n * p - context
    xd = x.double()
    loc = xd.mean(-1, keepdim=True)
    scale = ((xd - loc) ** 2).mean(-1, keepdim=True).sqrt()
    scale = torch.where(scale == 0, torch.full_like(scale, 1e-5), scale)
    scaled = torch.arcsinh((xd - loc) / scale)
    cm = (~mask).double()
    scaled = torch.cat([torch.full((b, pad), float("nan"), dtype=torch.float64), scaled], 1).reshape(b, n, p)
    pmask = torch.nan_to_num(torch.cat([torch.full((b, pad), float("nan"), dtype=torch.float64), cm], 1), nan=0.0).reshape(b, n, p)
    pctx = torch.where(pmask > 0, scaled, torch.zeros_like(scaled))
    tenc = (torch.arange(-n * p, 0, dtype=torch.float32) / 8192.0).double().reshape(1, n, p).expand(b, -1, -1)
    ref = torch.cat([tenc, pctx, pmask], -1).float()
    ref_am = pmask.sum(-1) > 0
    patched, am, gloc, gscale = ops.chronos2_patchify_norm(x.to(DEV), mask.to(DEV), out_cols=out_cols)
    assert torch.equal(am.cpu(), ref_am)
    assert (gloc.cpu() - loc[:, 0].float()).abs().max().item() < 1e-5
    assert _rel(gscale.cpu()[[0, 1, 2] + list(range(4, b))], scale[:, 0].float()[[0, 1, 2] + list(range(4, b))]) < 1e-5
    got = patched.cpu().reshape(b, n, out_cols)
    ok = [i for i in range(b) if i != 3]
    assert (got[ok, :, :48] - ref[ok]).abs().max().item() < 2e-5
    if out_cols > 48:
        assert got[..., 48:].abs().max().item() == 0.0


# ----------------------------------------------------------------------------- row norms
@pytest.mark.parametrize("cols", [1280, 768])
@pytest.mark.parametrize("out_dtype", [DT_F32, DT_BF16, DT_BF16_SPLIT])
def test_rmsnorm(cols, out_dtype):
    rows = 1003
    gen = torch.Generator(device=DEV).manual_seed(cols)
    x = torch.randn(rows, cols, generator=gen, device=DEV) * 2
    w = 1 + 0.1 * torch.randn(cols, generator=gen, device=DEV)
    out = ops.rmsnorm(x, w, 1e-6, out_dtype)
    xd = x.double()
    ref = (w.double() * (xd * torch.rsqrt(xd.pow(2).mean(-1, keepdim=True) + 1e-6))).float()
    tol = 6e-3 if out_dtype == DT_BF16 else 3e-5
    assert _rel(_to_float(out, out_dtype), ref) < tol


@pytest.mark.parametrize("a_bf16", [False, True])
@pytest.mark.parametrize("with_post", [True, False])
@pytest.mark.parametrize("with_next", [True, False])
def test_norm_residual_norm(a_bf16, with_post, with_next):
    rows, cols = 517, 1280
    gen = torch.Generator(device=DEV).manual_seed(9)
    a = torch.randn(rows, cols, generator=gen, device=DEV)
    if a_bf16:
        a = a.to(torch.bfloat16)
    x = torch.randn(rows, cols, generator=gen, device=DEV)
    w1 = 1 + 0.1 * torch.randn(cols, generator=gen, device=DEV)
    w2 = 1 + 0.1 * torch.randn(cols, generator=gen, device=DEV)
    y = torch.empty(rows, cols, device=DEV)
    yn = ops.alloc(rows, cols, DT_BF16_SPLIT, torch.device(DEV))
    ops.norm_residual_norm(a, x, w1 if with_post else None, w2 if with_next else None, 1e-6, y, DT_BF16_SPLIT, yn)
    ad = a.double()
    if with_post:
        ad = w1.double() * (ad * torch.rsqrt(ad.pow(2).mean(-1, keepdim=True) + 1e-6))
    ry = ad + x.double()
    ryn = w2.double() * (ry * torch.rsqrt(ry.pow(2).mean(-1, keepdim=True) + 1e-6)) if with_next else ry
    assert _rel(y, ry.float()) < 3e-6
    assert _rel(ops.split_to_float(yn), ryn.float()) < 3e-5
    # in-place residual update (y aliases x) is what the layer loop does
    x2 = x.clone()
    ops.norm_residual_norm(a, x2, w1 if with_post else None, w2 if with_next else None, 1e-6, x2, DT_BF16_SPLIT, yn)
    assert torch.equal(x2, y)


# ----------------------------------------------------------------------------- attention
def _ref_attention(qkv, b, n, h, hd, patch_mask, inv_freq, qw, kw, per_dim):
    d = h * hd
    q, k, v = qkv.double().reshape(b, n, 3, h, hd).unbind(2)
    nm = patch_mask.sum(-1)
    pos = (torch.arange(n, device=qkv.device)[None, :] - nm[:, None]).float()
    freqs = (pos[..., None] * inv_freq[None, None, :]).float()
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos().double()[:, :, None, :], emb.sin().double()[:, :, None, :]

    def rot(t):
        return torch.cat([-t[..., hd // 2 :], t[..., : hd // 2]], -1)

    q = q * cos + rot(q) * sin
    k = k * cos + rot(k) * sin
    q = qw.double() * (q * torch.rsqrt(q.pow(2).mean(-1, keepdim=True) + 1e-6))
    k = kw.double() * (k * torch.rsqrt(k.pow(2).mean(-1, keepdim=True) + 1e-6))
    q = q * (torch.nn.functional.softplus(per_dim.double()) * (1.442695041 / math.sqrt(hd)))
    s = torch.einsum("bqhd,bkhd->bhqk", q, k)
    causal = torch.tril(torch.ones(n, n, dtype=torch.bool, device=qkv.device))
    allowed = causal[None, None] & (~patch_mask)[:, None, None, :]
    s = s + torch.where(allowed, 0.0, torch.finfo(torch.float32).min).double()
    p = torch.softmax(s.float(), -1).double()
    o = torch.einsum("bhqk,bkhd->bqhd", p, v)
    return o.reshape(b * n, d).float()


@pytest.mark.parametrize("n", [16, 64, 5])
@pytest.mark.parametrize("qkv_bf16", [False, True])
def test_timesfm_attention(n, qkv_bf16):
    b, h, hd = 9, 16, 80
    gen = torch.Generator(device=DEV).manual_seed(n)
    qkv = torch.randn(b * n, 3 * h * hd, generator=gen, device=DEV)
    if qkv_bf16:
        qkv = qkv.to(torch.bfloat16)
    pm = torch.zeros(b, n, dtype=torch.bool, device=DEV)
    pm[1, : n // 2] = True
    pm[2, :1] = True
    pm[3, :] = True
    nm = pm.sum(-1).int()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)).to(DEV)
    qw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    kw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    per_dim = 0.5 * torch.randn(hd, generator=gen, device=DEV)
    q_scale = (torch.nn.functional.softplus(per_dim) * (1.442695041 / math.sqrt(hd))).contiguous()
    out = ops.timesfm_attention(qkv, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_F32)
    ref = _ref_attention(qkv.float(), b, n, h, hd, pm, inv_freq, qw, kw, per_dim)
    assert _rel(out, ref) < 2e-5
    out_s = ops.timesfm_attention(qkv, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16_SPLIT)
    assert _rel(ops.split_to_float(out_s), ref) < 5e-5


@pytest.mark.parametrize("n", [16, 64, 5, 32, 40, 1])
def test_timesfm_attention_tensor_core_path(n):
    """bf16 in / bf16 out goes through the mma.sync kernel; the fp32 SIMT kernel is the yardstick as well."""
    b, h, hd = 37, 16, 80
    gen = torch.Generator(device=DEV).manual_seed(100 + n)
    qkv = torch.randn(b * n, 3 * h * hd, generator=gen, device=DEV).to(torch.bfloat16)
    pm = torch.zeros(b, n, dtype=torch.bool, device=DEV)
    pm[1, : n // 2] = True
    pm[2, :1] = True
    pm[3, :] = True
    pm[4, : n - 1] = True
    nm = pm.sum(-1).int()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)).to(DEV)
    qw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    kw = 1 + 0.1 * torch.randn(hd, generator=gen, device=DEV)
    per_dim = 0.5 * torch.randn(hd, generator=gen, device=DEV)
    q_scale = (torch.nn.functional.softplus(per_dim) * (1.442695041 / math.sqrt(hd))).contiguous()
    ref = _ref_attention(qkv.float(), b, n, h, hd, pm, inv_freq, qw, kw, per_dim)
    lib = _lib.load()
    c0 = _lib.launch_count()
    out = ops.timesfm_attention(qkv, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16)
    assert _lib.launch_count() == c0 + 1
    _lib.check(lib.tsfmx_attention_force_simt(1))
    try:
        simt = ops.timesfm_attention(qkv, b, n, h, hd, pm, nm, inv_freq, qw, kw, q_scale, 1e-6, DT_BF16)
    finally:
        _lib.check(lib.tsfmx_attention_force_simt(0))
    assert not torch.isnan(out.float()).any()
    assert _rel(simt.float(), ref) < 6e-3
    assert _rel(out.float(), ref) < 2e-2, _rel(out.float(), ref)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-2


# ----------------------------------------------------------------------------- cluster-fused GEMM + row norm
@pytest.mark.parametrize("precision", [PREC_BF16, PREC_BF16X3])
@pytest.mark.parametrize("n,k,m", [(1280, 1280, 1000), (768, 3072, 300), (1280, 1280, 128 * 40 + 5), (768, 768, 64)])
@pytest.mark.parametrize("post,nxt", [(True, True), (False, True), (True, False)])
def test_gemm_rownorm(precision, n, k, m, post, nxt):
    gen = torch.Generator(device=DEV).manual_seed(n + k + m)
    a, af = _make_operand(m, k, precision, gen)
    b32 = torch.randn(n, k, generator=gen, device=DEV) / math.sqrt(k)
    if precision == PREC_BF16:
        b = b32.to(torch.bfloat16)
        bf = b.float()
    else:
        b, bf = ops.cast_rows(b32, DT_BF16_SPLIT), b32
    x = torch.randn(m, n, generator=gen, device=DEV)
    w1 = 1 + 0.1 * torch.randn(n, generator=gen, device=DEV)
    w2 = 1 + 0.1 * torch.randn(n, generator=gen, device=DEV)
    adt = DT_BF16_SPLIT if precision == PREC_BF16X3 else DT_BF16
    y = torch.full((m, n), float("nan"), device=DEV)
    yn = ops.alloc(m, n, adt, torch.device(DEV))
    ops.gemm_rownorm(a, b, k, m, n, precision, w1 if post else None, w2 if nxt else None, x, y, adt, yn, 1e-6)
    acc = af.double() @ bf.double().t()
    if post:
        acc = w1.double() * (acc * torch.rsqrt(acc.pow(2).mean(-1, keepdim=True) + 1e-6))
    ry = acc + x.double()
    ryn = w2.double() * (ry * torch.rsqrt(ry.pow(2).mean(-1, keepdim=True) + 1e-6)) if nxt else ry
    assert _rel(y, ry.float()) < 3e-5
    assert _rel(_to_float(yn, adt), ryn.float()) < (6e-3 if adt == DT_BF16 else 5e-5)
    # in place (y aliases x), the way the layer loop uses it
    x2 = x.clone()
    ops.gemm_rownorm(a, b, k, m, n, precision, w1 if post else None, w2 if nxt else None, x2, x2, adt, yn, 1e-6)
    assert torch.equal(x2, y)

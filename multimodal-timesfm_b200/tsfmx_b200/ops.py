"""Thin tensor-level wrappers over the C ABI: one function per entry point of ``include/tsfmx_b200.h``.

PyTorch is only the allocator / stream provider here; all arithmetic happens in the CUDA library.
Storage conventions: ``DT_BF16`` -> ``torch.bfloat16 [R, K]``; ``DT_BF16_SPLIT`` -> ``torch.bfloat16 [R, 2K]``
(hi | lo); ``DT_F32`` -> ``torch.float32 [R, K]``.
"""

from __future__ import annotations

import ctypes
import functools
from collections.abc import Sequence

import torch

from . import _lib
from ._lib import ACT_NONE, DT_BF16, DT_BF16_SPLIT, DT_F32, PREC_BF16, PREC_BF16X3, GemmArgs, check, ptr, stream


def _arg_tensors(args, kwargs):
    for a in (*args, *kwargs.values()):
        if isinstance(a, torch.Tensor):
            yield a
        elif isinstance(a, (list, tuple)):  # gemm segments: [(A, B, k), ...]
            for item in a:
                if isinstance(item, torch.Tensor):
                    yield item
                elif isinstance(item, (list, tuple)):
                    for t in item:
                        if isinstance(t, torch.Tensor):
                            yield t


def _on_operand_device(fn):
    """Run an entry point on the device that holds its operands.

    The C ABI takes raw pointers and a stream; kernels, TMA descriptors and the SM count all belong to the CURRENT
    device.  A caller that built its model on ``cuda:1`` without ``torch.cuda.set_device(1)`` would otherwise launch on
    GPU 0's stream with GPU 1's pointers (illegal address, or silent execution on the wrong GPU with peer access on).
    Operands on different devices raise ``TsfmxError``."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        index = None
        for t in _arg_tensors(args, kwargs):
            if not t.is_cuda:
                continue
            if index is None:
                index = t.device.index
            elif t.device.index != index:
                raise _lib.TsfmxError(
                    f"{fn.__name__}: operands live on different devices (cuda:{index} and cuda:{t.device.index})"
                )
        if index is None or index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(index):
            return fn(*args, **kwargs)

    return wrapper


def act_dtype(precision: int) -> int:
    """Storage of a GEMM input activation for a precision mode."""
    return DT_BF16_SPLIT if precision == PREC_BF16X3 else DT_BF16


def alloc(rows: int, cols: int, dtype: int, device: torch.device) -> torch.Tensor:
    if dtype == DT_F32:
        return torch.empty(rows, cols, dtype=torch.float32, device=device)
    if dtype == DT_BF16:
        return torch.empty(rows, cols, dtype=torch.bfloat16, device=device)
    return torch.empty(rows, 2 * cols, dtype=torch.bfloat16, device=device)


def _as_u8(mask: torch.Tensor) -> torch.Tensor:
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    if mask.dtype == torch.uint8:
        return mask.contiguous()
    return (mask != 0).contiguous().view(torch.uint8)


@_on_operand_device
def timesfm_patchify_norm(
    x: torch.Tensor, mask: torch.Tensor, patch_len: int = 32, tokens_dtype: int = DT_F32
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> tokens [B*N, 2P], mu [B, N], sigma [B, N], patch_mask [B, N] (bool), num_masked [B] (int32)."""
    _lib.require_cuda(x, mask)
    lib = _lib.load()
    x = x.contiguous().float()
    m8 = _as_u8(mask)
    b, c = x.shape
    n = c // patch_len
    tokens = alloc(b * n, 2 * patch_len, tokens_dtype, x.device)
    mu = torch.empty(b, n, dtype=torch.float32, device=x.device)
    sigma = torch.empty(b, n, dtype=torch.float32, device=x.device)
    pm = torch.empty(b, n, dtype=torch.uint8, device=x.device)
    nm = torch.empty(b, dtype=torch.int32, device=x.device)
    check(
        lib.tsfmx_timesfm_patchify_norm(
            ptr(x), ptr(m8), b, c, patch_len, tokens_dtype, ptr(tokens), ptr(mu), ptr(sigma), ptr(pm), ptr(nm), stream()
        )
    )
    return tokens, mu, sigma, pm.view(torch.bool), nm


@_on_operand_device
def chronos2_patchify_norm(
    x: torch.Tensor,
    mask: torch.Tensor,
    patch: int = 16,
    use_arcsinh: bool = True,
    time_encoding_scale: float = 8192.0,
    out_dtype: int = DT_F32,
    out_cols: int | None = None,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> patched [B*N, out_cols], attn_mask [B, N] (bool, True = observed), loc [B], scale [B]."""
    _lib.require_cuda(x, mask)
    lib = _lib.load()
    x = x.contiguous().float()
    m8 = _as_u8(mask)
    b, c = x.shape
    n = (c + patch - 1) // patch
    out_cols = 3 * patch if out_cols is None else out_cols
    patched = alloc(b * n, out_cols, out_dtype, x.device)
    am = torch.empty(b, n, dtype=torch.uint8, device=x.device)
    loc = torch.empty(b, dtype=torch.float32, device=x.device)
    scale = torch.empty(b, dtype=torch.float32, device=x.device)
    check(
        lib.tsfmx_chronos2_patchify_norm(
            ptr(x), ptr(m8), b, c, patch, int(use_arcsinh), float(time_encoding_scale), out_dtype, out_cols,
            ptr(patched), ptr(am), ptr(loc), ptr(scale), stream(),
        )
    )
    return patched, am.view(torch.bool), loc, scale


@_on_operand_device
def chronos_t5_tokenize(
    x: torch.Tensor,
    boundaries: torch.Tensor,
    n_special: int = 2,
    n_tokens: int = 4096,
    pad_id: int = 0,
    eos_id: int = 1,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> ids [B, C+1] int64, attention_mask [B, C+1] bool, scale [B] fp32."""
    _lib.require_cuda(x, boundaries)
    lib = _lib.load()
    x = x.contiguous().float()
    boundaries = boundaries.contiguous().float()
    b, c = x.shape
    ids = torch.empty(b, c + 1, dtype=torch.int64, device=x.device)
    am = torch.empty(b, c + 1, dtype=torch.uint8, device=x.device)
    scale = torch.empty(b, dtype=torch.float32, device=x.device)
    check(
        lib.tsfmx_chronos_t5_tokenize(
            ptr(x), b, c, ptr(boundaries), boundaries.numel(), n_special, n_tokens, pad_id, eos_id, ptr(ids), ptr(am),
            ptr(scale), stream(),
        )
    )
    return ids, am.view(torch.bool), scale


@_on_operand_device
def chronos_t5_dequantize(ids: torch.Tensor, centers: torch.Tensor, scale: torch.Tensor, n_special: int = 2) -> torch.Tensor:
    _lib.require_cuda(ids, centers, scale)
    lib = _lib.load()
    ids = ids.contiguous()
    centers = centers.contiguous().float()
    scale = scale.contiguous().float()
    b, length = ids.shape
    out = torch.empty(b, length, dtype=torch.float32, device=ids.device)
    check(
        lib.tsfmx_chronos_t5_dequantize(
            ptr(ids), b, length, ptr(centers), centers.numel(), n_special, ptr(scale), ptr(out), stream()
        )
    )
    return out


@_on_operand_device
def cast_rows(x: torch.Tensor, out_dtype: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 [R, K] (row stride allowed) -> bf16 [R, K] or split bf16 [R, 2K]."""
    _lib.require_cuda(x)
    lib = _lib.load()
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = alloc(rows, cols, out_dtype, x.device)
    check(lib.tsfmx_cast_rows(ptr(x), rows, cols, x.stride(0), out_dtype, ptr(out), stream()))
    return out


@_on_operand_device
def gemm(
    segments: Sequence[tuple[torch.Tensor, torch.Tensor, int]],
    m: int,
    n: int,
    out: torch.Tensor,
    d_dtype: int,
    precision: int = PREC_BF16,
    act: int = ACT_NONE,
    bias: torch.Tensor | None = None,
    row_scale: torch.Tensor | None = None,
    row_shift: torch.Tensor | None = None,
    residual: torch.Tensor | None = None,
    n_store: int = 0,
    ldd: int | None = None,
    split_off: int = 0,
    aux: torch.Tensor | None = None,
    pre_act: torch.Tensor | None = None,
) -> torch.Tensor:
    """``out = epilogue(sum_s A_s @ B_s^T)``; each segment is (A [m, k or 2k], B [n, k or 2k], k).

    ``aux``: saved activation input for the ``*_GRAD`` epilogues; ``pre_act``: optional second output receiving
    ``acc + bias`` before the activation (fp32 or bf16)."""
    lib = _lib.load()
    args = GemmArgs()
    args.m, args.n, args.num_segments = m, n, len(segments)
    for i, (a, b, k) in enumerate(segments):
        _lib.require_cuda(a, b)
        assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
        assert a.stride(-1) == 1 and b.stride(-1) == 1
        args.seg[i].a, args.seg[i].lda = a.data_ptr(), a.stride(0)
        args.seg[i].b, args.seg[i].ldb = b.data_ptr(), b.stride(0)
        args.seg[i].k = k
    args.precision, args.act = precision, act
    args.bias, args.row_scale, args.row_shift = ptr(bias), ptr(row_scale), ptr(row_shift)
    args.residual = ptr(residual)
    args.ldr = residual.stride(0) if residual is not None else 0
    args.d = out.data_ptr()
    args.ldd = out.stride(0) if ldd is None else ldd
    args.d_dtype, args.n_store, args.split_off = d_dtype, n_store, split_off
    if aux is not None:
        args.aux, args.ld_aux = aux.data_ptr(), aux.stride(0)
        args.aux_dtype = DT_BF16 if aux.dtype == torch.bfloat16 else DT_F32
    if pre_act is not None:
        args.pre_act, args.ld_pre = pre_act.data_ptr(), pre_act.stride(0)
        args.pre_act_dtype = DT_BF16 if pre_act.dtype == torch.bfloat16 else DT_F32
    check(lib.tsfmx_gemm(ctypes.byref(args), stream()))
    return out


@_on_operand_device
def gemm_rownorm(
    a: torch.Tensor,
    b: torch.Tensor,
    k: int,
    m: int,
    n: int,
    precision: int,
    w_post: torch.Tensor | None,
    w_next: torch.Tensor | None,
    x: torch.Tensor,
    y: torch.Tensor,
    yn_dtype: int,
    yn: torch.Tensor | None,
    eps: float,
) -> None:
    """``y = rmsnorm(a @ b^T) * w_post + x`` (or ``a @ b^T + x``) and ``yn = rmsnorm(y) * w_next`` (or ``y``) in one launch."""
    lib = _lib.load()
    _lib.require_cuda(a, b, x, y)
    seg = _lib.GemmSegment()
    seg.a, seg.lda, seg.b, seg.ldb, seg.k = a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), k
    check(
        lib.tsfmx_gemm_rownorm(
            ctypes.byref(seg), m, n, precision, ptr(w_post), ptr(w_next), ptr(x), ptr(y), yn_dtype, ptr(yn), eps, stream()
        )
    )


@_on_operand_device
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out_dtype: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _lib.require_cuda(x, w)
    lib = _lib.load()
    rows, cols = x.shape
    if out is None:
        out = alloc(rows, cols, out_dtype, x.device)
    check(lib.tsfmx_rmsnorm(ptr(x), rows, cols, ptr(w), eps, out_dtype, ptr(out), stream()))
    return out


@_on_operand_device
def norm_residual_norm(
    a: torch.Tensor,
    x: torch.Tensor,
    w_post: torch.Tensor | None,
    w_next: torch.Tensor | None,
    eps: float,
    y: torch.Tensor | None,
    yn_dtype: int,
    yn: torch.Tensor | None,
) -> None:
    _lib.require_cuda(a, x)
    lib = _lib.load()
    rows, cols = x.shape
    a_dtype = DT_BF16 if a.dtype == torch.bfloat16 else DT_F32
    check(
        lib.tsfmx_norm_residual_norm(
            ptr(a), a_dtype, ptr(x), rows, cols, ptr(w_post), ptr(w_next), eps, ptr(y), yn_dtype, ptr(yn), stream()
        )
    )


@_on_operand_device
def timesfm_attention(
    qkv: torch.Tensor,
    batch: int,
    num_patches: int,
    num_heads: int,
    head_dim: int,
    patch_mask: torch.Tensor | None,
    num_masked: torch.Tensor | None,
    inv_freq: torch.Tensor,
    q_ln_w: torch.Tensor,
    k_ln_w: torch.Tensor,
    q_scale: torch.Tensor,
    eps: float,
    out_dtype: int,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    _lib.require_cuda(qkv)
    lib = _lib.load()
    qkv_dtype = DT_BF16 if qkv.dtype == torch.bfloat16 else DT_F32
    if out is None:
        out = alloc(batch * num_patches, num_heads * head_dim, out_dtype, qkv.device)
    pm = None if patch_mask is None else _as_u8(patch_mask)
    check(
        lib.tsfmx_timesfm_attention(
            ptr(qkv), qkv_dtype, batch, num_patches, num_heads, head_dim, ptr(pm), ptr(num_masked), ptr(inv_freq),
            ptr(q_ln_w), ptr(k_ln_w), ptr(q_scale), eps, out_dtype, ptr(out), stream(),
        )
    )
    return out


def split_to_float(t: torch.Tensor) -> torch.Tensor:
    """Debug helper: split bf16 [R, 2K] -> fp32 [R, K] (hi + lo)."""
    k = t.shape[1] // 2
    return t[:, :k].float() + t[:, k:].float()


# ----------------------------------------------------------------------------- backward pass
def _dt(t: torch.Tensor) -> int:
    return DT_BF16 if t.dtype == torch.bfloat16 else DT_F32


@_on_operand_device
def rmsnorm_bwd_chain(
    g_res: torch.Tensor | None,
    v1: torch.Tensor | None,
    w1: torch.Tensor | None,
    g1: torch.Tensor | None,
    v2: torch.Tensor | None,
    w2: torch.Tensor | None,
    eps: float,
    g_total: torch.Tensor | None,
    g2_dtype: int,
    g2: torch.Tensor | None,
    rows: int,
    cols: int,
    dw1: torch.Tensor | None = None,
    dw2: torch.Tensor | None = None,
) -> None:
    """``dw1`` / ``dw2`` (full fine-tuning): fp32 [cols], ACCUMULATED into - the gradients of the scales ``w1`` / ``w2``
    from the same pass over the rows."""
    lib = _lib.load()
    for t in (dw1, dw2):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == cols)
    check(
        lib.tsfmx_rmsnorm_bwd_chain_wgrad(
            ptr(g_res), ptr(v1), _dt(v1) if v1 is not None else 0, ptr(w1), ptr(g1), _dt(g1) if g1 is not None else 0,
            ptr(v2), _dt(v2) if v2 is not None else 0, ptr(w2), rows, cols, eps, ptr(g_total), g2_dtype, ptr(g2),
            ptr(dw1), ptr(dw2), stream(),
        )
    )


@_on_operand_device
def colsum_wgrad(g: torch.Tensor, v: torch.Tensor | None = None, eps: float = 0.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """``out[c] += sum_r g[r, c] * v_hat[r, c]`` (RMSNorm scale gradient, ``v_hat`` = RMS-normalised ``v``) or, without
    ``v``, ``out[c] += sum_r g[r, c]`` (bias gradient).  ``out`` fp32 [cols] is created zeroed when not given."""
    lib = _lib.load()
    _lib.require_cuda(g, v)
    rows, cols = g.shape
    if out is None:
        out = torch.zeros(cols, dtype=torch.float32, device=g.device)
    check(lib.tsfmx_colsum_wgrad(ptr(v), _dt(v) if v is not None else 0, ptr(g), _dt(g), rows, cols, eps,
                                 int(v is not None), ptr(out), stream()))
    return out


@_on_operand_device
def timesfm_attention_bwd(
    qkv: torch.Tensor,
    d_out: torch.Tensor,
    batch: int,
    num_patches: int,
    num_heads: int,
    head_dim: int,
    patch_mask: torch.Tensor | None,
    num_masked: torch.Tensor | None,
    inv_freq: torch.Tensor,
    q_ln_w: torch.Tensor,
    k_ln_w: torch.Tensor,
    q_scale: torch.Tensor,
    eps: float,
    dqkv_dtype: int,
    dqkv: torch.Tensor | None = None,
    dparams: torch.Tensor | None = None,
) -> torch.Tensor:
    """``dparams``: optional fp32 [2 * head_dim] accumulator for d/d(q_ln_w * q_scale) and d/d(k_ln_w)."""
    lib = _lib.load()
    _lib.require_cuda(qkv, d_out)
    if dqkv is None:
        dqkv = alloc(batch * num_patches, 3 * num_heads * head_dim, dqkv_dtype, qkv.device)
    pm = None if patch_mask is None else _as_u8(patch_mask)
    check(
        lib.tsfmx_timesfm_attention_bwd(
            ptr(qkv), _dt(qkv), ptr(d_out), _dt(d_out), batch, num_patches, num_heads, head_dim, ptr(pm),
            ptr(num_masked), ptr(inv_freq), ptr(q_ln_w), ptr(k_ln_w), ptr(q_scale), eps, dqkv_dtype, ptr(dqkv),
            ptr(dparams), stream(),
        )
    )
    return dqkv


def _storage_dtype(t: torch.Tensor, logical_cols: int) -> int:
    if t.dtype == torch.float32:
        return DT_F32
    return DT_BF16_SPLIT if t.shape[1] == 2 * logical_cols else DT_BF16


@_on_operand_device
def transpose_mask(
    x: torch.Tensor, rows: int, cols: int, out_dtype: int, mask: torch.Tensor | None = None
) -> tuple[torch.Tensor, int]:
    """[rows, cols] (f32 / bf16 / split) -> K-major transposed [cols, kpad] (bf16 / split), kpad = rows rounded up to 64.
    Returns (out, kpad)."""
    lib = _lib.load()
    _lib.require_cuda(x)
    kpad = (rows + 63) // 64 * 64
    out = alloc(cols, kpad, out_dtype, x.device)
    check(
        lib.tsfmx_transpose_mask(
            ptr(x), _storage_dtype(x, cols), rows, cols, x.stride(0), ptr(mask),
            _storage_dtype(mask, cols) if mask is not None else 0, mask.stride(0) if mask is not None else 0,
            out_dtype, ptr(out), kpad, stream(),
        )
    )
    return out, kpad


@_on_operand_device
def mask_cast_rows(x: torch.Tensor, mask: torch.Tensor, out_dtype: int) -> torch.Tensor:
    """fp32 [R, C] gated by mask[r, c] > 0 -> f32 / bf16 / split [R, C]."""
    lib = _lib.load()
    _lib.require_cuda(x, mask)
    rows, cols = x.shape
    out = alloc(rows, cols, out_dtype, x.device)
    check(
        lib.tsfmx_mask_cast_rows(
            ptr(x), rows, cols, ptr(mask), _storage_dtype(mask, cols), mask.stride(0), out_dtype, ptr(out), stream()
        )
    )
    return out


# ----------------------------------------------------------------------------- whole-stack entry point
def timesfm_stack_table(layers: Sequence[dict], inv_freq: torch.Tensor, model_dims: int, num_heads: int, head_dim: int,
                        ff_dims: int, precision: int, eps: float):
    """Pack the per-layer device pointers into the C ABI's ``tsfmx_timesfm_stack``.  The returned object keeps the
    ctypes arrays (and the tensors they point at) alive; build it once per set of packed weights."""
    arr = (_lib.TimesfmLayer * len(layers))()
    for dst, lw in zip(arr, layers):
        dst.qkv, dst.out, dst.ff0, dst.ff1 = (lw[k].data_ptr() for k in ("qkv", "out", "ff0", "ff1"))
        dst.pre_attn_ln, dst.post_attn_ln = lw["pre_attn"].data_ptr(), lw["post_attn"].data_ptr()
        dst.pre_ff_ln, dst.post_ff_ln = lw["pre_ff"].data_ptr(), lw["post_ff"].data_ptr()
        dst.q_ln, dst.k_ln, dst.q_scale = lw["q_ln"].data_ptr(), lw["k_ln"].data_ptr(), lw["q_scale"].data_ptr()
    table = _lib.TimesfmStack()
    table.num_layers, table.model_dims, table.num_heads, table.head_dim = len(layers), model_dims, num_heads, head_dim
    table.ff_dims, table.precision, table.eps = ff_dims, precision, eps
    table.inv_freq = inv_freq.data_ptr()
    table.layers = ctypes.cast(arr, ctypes.POINTER(_lib.TimesfmLayer))
    table._keepalive = (arr, list(layers), inv_freq)
    return table


@_on_operand_device
def timesfm_stack_fwd(table, x: torch.Tensor, batch: int, num_patches: int, patch_mask: torch.Tensor | None,
                      num_masked: torch.Tensor | None) -> torch.Tensor:
    """All decoder layers in one library call: x [B * N, D] fp32 -> y [B * N, D] fp32 (``tsfmx_timesfm_stack_fwd``)."""
    lib = _lib.load()
    _lib.require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    need = int(lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), batch, num_patches))
    if need == 0 and batch > 0:
        raise _lib.TsfmxError("tsfmx_timesfm_stack_workspace_bytes rejected the weight table")
    workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=x.device)
    y = torch.empty_like(x)
    pm = None if patch_mask is None else _as_u8(patch_mask)
    check(lib.tsfmx_timesfm_stack_fwd(ctypes.byref(table), batch, num_patches, ptr(x), ptr(pm), ptr(num_masked),
                                      ptr(workspace), need, ptr(y), stream()))
    return y


def _token_major_ok(t: torch.Tensor, cols: int) -> bool:
    return (t.dtype == torch.bfloat16 and t.dim() == 2 and t.shape[1] == cols and t.stride(1) == 1
            and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0)


@_on_operand_device
def wgrad(dy: torch.Tensor, x: torch.Tensor, rows: int, n_out: int, k_in: int, precision: int) -> torch.Tensor:
    """Weight gradient of ``y = x W^T``: dW [n_out, k_in] = dY^T X, a tcgen05 GEMM with K = rows (split-K inside the GEMM
    when the tile count is small).  bf16 operands are consumed where they lie, [tokens, features] row-major = MN-major
    for this product (``tsfmx_gemm_wgrad``); fp32 / split operands (parity mode, fp32 head gradients) are transposed
    to K-major copies first."""
    if precision == PREC_BF16 and _token_major_ok(dy, n_out) and _token_major_ok(x, k_in) and n_out % 8 == 0 and k_in % 8 == 0:
        lib = _lib.load()
        _lib.require_cuda(dy, x)
        gw = torch.empty(n_out, k_in, dtype=torch.float32, device=dy.device)
        check(lib.tsfmx_gemm_wgrad(ptr(dy), dy.stride(0), ptr(x), x.stride(0), rows, n_out, k_in, ptr(gw), k_in, stream()))
        return gw
    adt = act_dtype(precision)
    dy_t, kpad = transpose_mask(dy, rows, n_out, adt)
    x_t, _ = transpose_mask(x, rows, k_in, adt)
    gw = torch.empty(n_out, k_in, dtype=torch.float32, device=dy.device)
    gemm([(dy_t, x_t, kpad)], n_out, k_in, gw, DT_F32, precision=precision)
    return gw


# ----------------------------------------------------------------------------- TimesFM AR decode / forecast extras
@_on_operand_device
def timesfm_patchify_continue(
    values: torch.Tensor, state: tuple[torch.Tensor, torch.Tensor, torch.Tensor], patches: int, patch_len: int = 32,
    tokens_dtype: int = DT_F32,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The ``patches`` new input patches of a decode step.  ``values`` is a [B, patches * patch_len] fp32 VIEW (any
    strides: a channel of the [B, steps, Q] forecast buffer); ``state`` = running (n, mu, sigma), each [B] fp32,
    UPDATED IN PLACE.  -> tokens [B * patches, 2 * patch_len], mu [B, patches], sigma [B, patches]."""
    _lib.require_cuda(values, *state)
    lib = _lib.load()
    assert values.dtype == torch.float32 and values.dim() == 2 and values.shape[1] == patches * patch_len
    b = values.shape[0]
    n, mu_s, sigma_s = state
    for t in state:
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == b
    tokens = alloc(b * patches, 2 * patch_len, tokens_dtype, values.device)
    mu = torch.empty(b, patches, dtype=torch.float32, device=values.device)
    sigma = torch.empty(b, patches, dtype=torch.float32, device=values.device)
    check(
        lib.tsfmx_timesfm_patchify_continue(
            ptr(values), values.stride(0), values.stride(1), b, patches, patch_len, ptr(n), ptr(mu_s), ptr(sigma_s),
            tokens_dtype, ptr(tokens), ptr(mu), ptr(sigma), stream(),
        )
    )
    return tokens, mu, sigma


@_on_operand_device
def timesfm_attention_decode(
    regions: Sequence[torch.Tensor],
    batch: int,
    num_heads: int,
    head_dim: int,
    patch_mask: torch.Tensor | None,
    num_masked: torch.Tensor | None,
    inv_freq: torch.Tensor,
    q_ln_w: torch.Tensor,
    k_ln_w: torch.Tensor,
    q_scale: torch.Tensor,
    eps: float,
    out_dtype: int,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """Attention of the LAST region's tokens (the new patches of a decode step) against every token of every region.
    ``regions``: raw qkv matrices [B * tokens_r, 3 * H * hd] (the prefill's first, then one per decode step) - the KV
    cache is whatever the qkv GEMMs left in HBM."""
    lib = _lib.load()
    _lib.require_cuda(*regions)
    width = 3 * num_heads * head_dim
    n = len(regions)
    if n < 1 or n > _lib.MAX_KV_REGIONS:
        raise _lib.TsfmxError(f"timesfm_attention_decode: 1..{_lib.MAX_KV_REGIONS} regions, got {n}")
    dtypes = {r.dtype for r in regions}
    if len(dtypes) != 1:
        raise _lib.TsfmxError("timesfm_attention_decode: all regions must share one dtype")
    tokens = []
    for r in regions:
        assert r.dim() == 2 and r.shape[1] == width and r.is_contiguous() and r.shape[0] % batch == 0
        tokens.append(r.shape[0] // batch)
    ptrs = (ctypes.c_void_p * n)(*[r.data_ptr() for r in regions])
    toks = (ctypes.c_int32 * n)(*tokens)
    if out is None:
        out = alloc(batch * tokens[-1], num_heads * head_dim, out_dtype, regions[0].device)
    pm = None if patch_mask is None else _as_u8(patch_mask)
    n_ctx = 0 if pm is None else pm.shape[1]
    rope_len = (max(sum(tokens), n_ctx) + 63) // 64 * 64  # |position| < max(total tokens, padded prefix); cached per length
    table = rope_table(inv_freq, rope_len)
    check(
        lib.tsfmx_timesfm_attention_decode(
            ptrs, toks, n, _dt(regions[0]), batch, num_heads, head_dim, n_ctx, ptr(pm), ptr(num_masked), ptr(table),
            rope_len, ptr(inv_freq), ptr(q_ln_w), ptr(k_ln_w), ptr(q_scale), eps, out_dtype, ptr(out), stream(),
        )
    )
    return out


@_on_operand_device
def timesfm_forecast_finalize(
    pf: torch.Tensor, spread: torch.Tensor | None, inputs: torch.Tensor | None, batch: int, horizon: int,
    decode_index: int, flip: bool, use_continuous_quantile_head: bool, infer_is_positive: bool,
) -> torch.Tensor:
    """pf [(1 + flip) * B, steps, Q] (+ spread [(1 + flip) * B, spread_steps, Q]) -> forecast [B, horizon, Q]: flip
    combination, continuous quantile head, positivity clamp, horizon slice in one launch."""
    lib = _lib.load()
    _lib.require_cuda(pf, spread, inputs)
    assert pf.dtype == torch.float32 and pf.is_contiguous() and pf.shape[0] == (2 if flip else 1) * batch
    nq = pf.shape[2]
    out = torch.empty(batch, horizon, nq, dtype=torch.float32, device=pf.device)
    if spread is not None:
        assert spread.dtype == torch.float32 and spread.is_contiguous() and spread.shape[0] == pf.shape[0]
    if inputs is not None:
        inputs = inputs.contiguous().float()
    check(
        lib.tsfmx_timesfm_forecast_finalize(
            ptr(pf), ptr(spread), ptr(inputs), batch, 0 if inputs is None else inputs.shape[1], pf.shape[1],
            0 if spread is None else spread.shape[1], nq, horizon, decode_index, int(flip),
            int(use_continuous_quantile_head), int(infer_is_positive), ptr(out), stream(),
        )
    )
    return out


# ----------------------------------------------------------------------------- Chronos-2
_force_simt_encoder_attention = False  # test hook: run the fp32 SIMT kernel where the tensor-core one applies
_rope_tables: dict[tuple, torch.Tensor] = {}


@_on_operand_device
def rope_table(inv_freq: torch.Tensor, seq: int) -> torch.Tensor:
    """(cos, sin) of position * inv_freq for positions [0, seq): fp32 [seq, half, 2], cached per inv_freq tensor."""
    key = (inv_freq.data_ptr(), inv_freq._version, inv_freq.device, seq)
    table = _rope_tables.get(key)
    if table is None:
        if len(_rope_tables) > 64:
            _rope_tables.clear()
        table = torch.empty(seq, inv_freq.numel(), 2, dtype=torch.float32, device=inv_freq.device)
        check(_lib.load().tsfmx_rope_table(ptr(inv_freq), inv_freq.numel(), seq, ptr(table), stream()))
        _rope_tables[key] = table
    return table


@_on_operand_device
def encoder_attention(
    qkv: torch.Tensor,
    batch: int,
    seq: int,
    num_heads: int,
    head_dim: int,
    key_mask: torch.Tensor | None,
    inv_freq: torch.Tensor,
    out_dtype: int,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(qkv)
    if out is None:
        out = alloc(batch * seq, num_heads * head_dim, out_dtype, qkv.device)
    km = None if key_mask is None else _as_u8(key_mask)
    if (qkv.dtype == torch.bfloat16 and out_dtype == DT_BF16 and head_dim == 64 and seq <= 208
            and not _force_simt_encoder_attention):
        table = rope_table(inv_freq, seq)
        check(lib.tsfmx_encoder_attention_mma(ptr(qkv), batch, seq, num_heads, head_dim, ptr(km), ptr(table), ptr(out),
                                              stream()))
        return out
    check(
        lib.tsfmx_encoder_attention(
            ptr(qkv), _dt(qkv), batch, seq, num_heads, head_dim, ptr(km), ptr(inv_freq), out_dtype, ptr(out), stream()
        )
    )
    return out


@_on_operand_device
def encoder_attention_bwd(
    qkv: torch.Tensor, d_out: torch.Tensor, batch: int, seq: int, num_heads: int, head_dim: int,
    key_mask: torch.Tensor | None, inv_freq: torch.Tensor, dqkv_dtype: int, dqkv: torch.Tensor | None = None,
) -> torch.Tensor:
    """d_out [B*T, H*hd] -> dqkv [B*T, 3*H*hd] (split: [B*T, 6*H*hd]) of the Chronos-2 encoder attention core."""
    lib = _lib.load()
    _lib.require_cuda(qkv, d_out)
    if dqkv is None:
        dqkv = alloc(batch * seq, 3 * num_heads * head_dim, dqkv_dtype, qkv.device)
    km = None if key_mask is None else _as_u8(key_mask)
    table = rope_table(inv_freq, seq)
    stats = torch.empty(batch * num_heads * seq, 4, dtype=torch.float32, device=qkv.device)
    check(
        lib.tsfmx_encoder_attention_bwd(
            ptr(qkv), _dt(qkv), ptr(d_out), _dt(d_out), batch, seq, num_heads, head_dim, ptr(km), ptr(table), ptr(stats),
            dqkv_dtype, ptr(dqkv), stream(),
        )
    )
    return dqkv


@_on_operand_device
def chronos2_finalize(
    preds: torch.Tensor, batch: int, patches_used: int, num_quantiles: int, patch: int, horizon: int,
    use_arcsinh: bool, loc: torch.Tensor, scale: torch.Tensor,
) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(preds, loc, scale)
    out = torch.empty(batch, horizon, num_quantiles, dtype=torch.float32, device=preds.device)
    check(
        lib.tsfmx_chronos2_finalize(
            ptr(preds), batch, patches_used, num_quantiles, patch, horizon, int(use_arcsinh),
            ptr(loc.reshape(-1).contiguous()), ptr(scale.reshape(-1).contiguous()), ptr(out), stream(),
        )
    )
    return out


# ----------------------------------------------------------------------------- Chronos-T5
@_on_operand_device
def embed_rows(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """ids (any shape, int64) -> fp32 [ids.numel(), dims] rows of the fp32 ``table`` [vocab, dims]."""
    lib = _lib.load()
    _lib.require_cuda(ids, table)
    ids = ids.reshape(-1).contiguous().to(torch.int64)
    table = table.contiguous().float()
    out = torch.empty(ids.numel(), table.shape[1], dtype=torch.float32, device=table.device)
    check(lib.tsfmx_embed_rows(ptr(ids), ids.numel(), table.shape[1], table.shape[0], ptr(table), ptr(out), stream()))
    return out


@_on_operand_device
def t5_sample_topk(logits: torch.Tensor, banned_id: int = -1, top_k: int = 1, temperature: float = 1.0,
                   uniform: torch.Tensor | None = None) -> torch.Tensor:
    """logits [rows, vocab] fp32 -> next ids [rows] int64: ban / temperature / top-k / softmax / inverse-CDF sampling in
    one launch (``top_k = 1``: greedy, first maximum)."""
    lib = _lib.load()
    _lib.require_cuda(logits, uniform)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.is_contiguous()
    rows, vocab = logits.shape
    if uniform is not None:
        uniform = uniform.reshape(-1).contiguous().float()
        assert uniform.numel() == rows
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    check(lib.tsfmx_t5_sample_topk(ptr(logits), rows, vocab, banned_id, float(temperature), int(top_k), ptr(uniform),
                                   ptr(out), stream()))
    return out


@_on_operand_device
def t5_attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    batch: int,
    tq: int,
    tk: int,
    num_heads: int,
    out_dtype: int,
    out: torch.Tensor,
    *,
    q_rows: tuple[int, int],
    kv_rows: tuple[int, int],
    out_rows: tuple[int, int],
    ldv: int | None = None,
    q_pos0: int = 0,
    causal: bool = False,
    key_mask: torch.Tensor | None = None,
    bias: torch.Tensor | None = None,
    bias_zero: int = 0,
    kv_batch_div: int = 1,
) -> torch.Tensor:
    """General T5 attention core (fp32 SIMT).  ``*_rows`` = (row stride, series stride) in elements of the q / k,v /
    out buffers; ``q``, ``k``, ``v`` may be views into one buffer (e.g. a [B, L, 3W] cache)."""
    lib = _lib.load()
    _lib.require_cuda(q, k, v, out)
    km = None if key_mask is None else _as_u8(key_mask)
    check(
        lib.tsfmx_t5_attention(
            ptr(q), _dt(q), q_rows[0], q_rows[1], ptr(k), ptr(v), _dt(k), kv_rows[0], kv_rows[0] if ldv is None else ldv,
            kv_rows[1], batch, tq, tk, num_heads, 64, q_pos0, int(causal), ptr(km), ptr(bias),
            0 if bias is None else bias.shape[-1], bias_zero, out_dtype, ptr(out), out_rows[0], out_rows[1], kv_batch_div,
            stream(),
        )
    )
    return out


@_on_operand_device
def t5_encoder_attention(
    qkv: torch.Tensor, batch: int, seq: int, num_heads: int, key_mask: torch.Tensor | None, bias: torch.Tensor | None,
    out_dtype: int, out: torch.Tensor | None = None,
) -> torch.Tensor:
    """Encoder self-attention over qkv [B*seq, 3*H*64]; ``bias`` fp32 [H, 2 seq - 1] indexed by (key - query) + seq - 1.
    bf16 in / bf16 out takes the tensor-core kernel, everything else the fp32 SIMT core."""
    lib = _lib.load()
    _lib.require_cuda(qkv)
    width = num_heads * 64
    if out is None:
        out = alloc(batch * seq, width, out_dtype, qkv.device)
    km = None if key_mask is None else _as_u8(key_mask)
    if qkv.dtype == torch.bfloat16 and out_dtype == DT_BF16 and seq <= 704 and not _force_simt_encoder_attention:
        check(lib.tsfmx_t5_encoder_attention_mma(ptr(qkv), batch, seq, num_heads, 64, ptr(km), ptr(bias), ptr(out), stream()))
        return out
    ld = 3 * width
    t5_attention(qkv, qkv[:, width:], qkv[:, 2 * width:], batch, seq, seq, num_heads, out_dtype, out,
                 q_rows=(ld, seq * ld), kv_rows=(ld, seq * ld), out_rows=(width, seq * width), key_mask=km, bias=bias,
                 bias_zero=seq - 1)
    return out

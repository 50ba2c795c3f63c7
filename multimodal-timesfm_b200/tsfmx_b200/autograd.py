"""Fusion fine-tune step on the B200 path (reference tsfmx/trainer.py:200-219).

In the reference's "multimodal" mode the adapter is frozen (trainer.py:76-77) and only
``MultimodalFusion`` is trained (trainer.py:119-123).  Fusion is injected at the *input* of the transformer
stack (decoder.py:66-71), so the loss gradient has to travel back through the head and every decoder layer
(activation gradients only) before it reaches the fusion projection, where the only weight gradients are
produced.  ``FusedForecastFunction`` is that whole pipeline as one ``torch.autograd.Function``: the forward
pass runs the same C-ABI kernels as inference while saving per-layer activations, the backward pass is
hand-written (dgrad GEMMs on pre-transposed frozen weights with SiLU'/ReLU' epilogues, fused RMSNorm/residual
backward, attention backward, fusion wgrad as a K = tokens tcgen05 GEMM).  PyTorch's autograd only sees the
function's inputs (the fusion weights) and its output (the forecast).
"""

from __future__ import annotations

import torch

from . import ops
from ._lib import ACT_RELU, ACT_RELU_GRAD, DT_F32, PRECISIONS, TsfmxError
from .fusion import _pad_k, _round64


class FusedForecastFunction(torch.autograd.Function):
    """forecast = decoder(horizon, inputs, masks, text); differentiable w.r.t. the fusion weights only."""

    @staticmethod
    def forward(ctx, decoder, horizon, inputs, masks, text, *fusion_weights):  # noqa: D401
        adapter, fusion = decoder.adapter, decoder.fusion
        precision = PRECISIONS[adapter.precision]
        pre = adapter.preprocess(inputs, masks)
        emb = pre.input_embeddings
        b, n, d = emb.shape
        fused, fusion_saved = fusion_forward_saving(fusion, emb.reshape(b * n, d), text.reshape(b * n, -1), precision)
        out_emb, stack_saved = adapter.forward_saving(fused.view(b, n, d), pre.masks)
        forecast, head_saved = adapter.postprocess_saving(horizon, out_emb, pre.normalization_stats)
        ctx.decoder = decoder
        ctx.saved = (fusion_saved, stack_saved, head_saved)
        ctx.shape = (b, n, d, horizon)
        ctx.num_weights = len(fusion_weights)
        return forecast

    @staticmethod
    def backward(ctx, grad_forecast):
        decoder = ctx.decoder
        adapter, fusion = decoder.adapter, decoder.fusion
        fusion_saved, stack_saved, head_saved = ctx.saved
        ctx.saved = None  # release the activations as early as possible
        b, n, d, horizon = ctx.shape
        precision = PRECISIONS[adapter.precision]
        # gradient w.r.t. the adapter's output embeddings (TimesFM: only the last patch feeds the head, reference
        # timesfm.py:129; Chronos-2: the first ceil(horizon / 16) forecast positions), then back through the stack
        d_out = adapter.postprocess_backward(head_saved, grad_forecast.contiguous().float())  # [B, N_out, D] fp32
        d_emb = adapter.forward_backward(stack_saved, d_out.reshape(-1, d))  # [B * N, D] fp32
        grads = fusion_backward(fusion, fusion_saved, d_emb, precision)
        return (None, None, None, None, None, *grads)


class FullFineTuneFunction(torch.autograd.Function):
    """forecast = decoder(horizon, inputs, masks, text or None); differentiable w.r.t. EVERY adapter parameter (and the
    fusion weights when text is given) - the reference's "baseline" mode unfreezes the adapter (trainer.py:78-79,123).
    Forward = the inference kernels + saved GEMM operands; backward = the activation-gradient chain of the fusion
    fine-tune plus one K = tokens weight-gradient GEMM per Linear, column reductions for norm scales / biases and the
    attention kernels' q_ln / k_ln / per-dim-scale gradients."""

    @staticmethod
    def forward(ctx, decoder, horizon, inputs, masks, text, names, num_fusion, *params):  # noqa: D401
        adapter, fusion = decoder.adapter, decoder.fusion
        precision = PRECISIONS[adapter.precision]
        pre, tok_saved = adapter.preprocess_saving(inputs, masks)
        emb = pre.input_embeddings
        b, n, d = emb.shape
        fusion_saved = None
        if text is not None:
            fused, fusion_saved = fusion_forward_saving(fusion, emb.reshape(b * n, d), text.reshape(b * n, -1), precision)
            emb = fused.view(b, n, d)
        out_emb, stack_saved = adapter.forward_saving(emb, pre.masks, for_wgrad=True)
        forecast, head_saved = adapter.postprocess_saving(horizon, out_emb, pre.normalization_stats)
        ctx.decoder, ctx.names, ctx.num_fusion = decoder, names, num_fusion
        ctx.saved = (tok_saved, fusion_saved, stack_saved, head_saved)
        ctx.shape = (b, n, d)
        return forecast

    @staticmethod
    def backward(ctx, grad_forecast):
        decoder = ctx.decoder
        adapter, fusion = decoder.adapter, decoder.fusion
        tok_saved, fusion_saved, stack_saved, head_saved = ctx.saved
        ctx.saved = None
        b, n, d = ctx.shape
        precision = PRECISIONS[adapter.precision]
        grads: dict[str, torch.Tensor] = {}
        # decoder.grad_ready_hook (set by a data-parallel trainer): called with every group of finished parameter
        # gradients, top of the network first, so that their all-reduce runs while the layers below are still in
        # their backward pass; hook.finish() joins the collectives before autograd sees the tensors
        hook = getattr(decoder, "grad_ready_hook", None)
        d_out = adapter.postprocess_backward(head_saved, grad_forecast.contiguous().float(), grads)
        if hook is not None:
            hook(list(grads.values()))
        d_emb = adapter.forward_backward(stack_saved, d_out.reshape(-1, d), grads, on_grads=hook)
        seen = set(grads)
        adapter.preprocess_backward(tok_saved, d_emb, grads)
        fusion_grads = [None] * ctx.num_fusion
        if fusion_saved is not None and ctx.num_fusion:
            fusion_grads = fusion_backward(fusion, fusion_saved, d_emb, precision)
        if hook is not None:
            hook([v for k, v in grads.items() if k not in seen] + [g for g in fusion_grads if g is not None])
            hook.finish()
        ordered = [grads.get(name) for name in ctx.names]
        return (None, None, None, None, None, None, None, *fusion_grads, *ordered)


def fusion_forward_saving(fusion, ts2: torch.Tensor, tx2: torch.Tensor, precision: int):
    """MultimodalFusion.forward (reference fusion.py:44-47) that also keeps what the backward pass needs."""
    adt = ops.act_dtype(precision)
    m = ts2.shape[0]
    dims = fusion.dims
    weights = fusion._packed_weights(precision)
    acts = [ops.cast_rows(_pad_k(tx2.float().contiguous()), adt)]  # h_0 = text (K padded to 64)
    pre_last = None
    out = None
    for i, w in enumerate(weights):
        n_out, k = dims[i + 1], _round64(dims[i])
        if i == len(weights) - 1:
            out = torch.empty(m, n_out, dtype=torch.float32, device=ts2.device)
            pre_last = torch.empty(m, n_out, dtype=torch.float32, device=ts2.device)
            ops.gemm([(acts[-1], w, k)], m, n_out, out, DT_F32, precision=precision, act=ACT_RELU,
                     residual=ts2.float().contiguous(), pre_act=pre_last)
        else:
            n_pad = _round64(n_out)
            nxt = ops.alloc(m, n_pad, adt, ts2.device)
            if n_pad != n_out:
                nxt.zero_()
            ops.gemm([(acts[-1], w, k)], m, n_out, nxt, adt, precision=precision, act=ACT_RELU, split_off=n_pad)
            acts.append(nxt)
    return out, (acts, pre_last)


def fusion_backward(fusion, saved, d_emb: torch.Tensor, precision: int) -> list[torch.Tensor]:
    """Weight gradients of every fusion Linear: dW_i = dpre_i^T h_{i-1}, dpre_{i-1} = (dpre_i W_i) * relu'(h_{i-1})."""
    acts, pre_last = saved
    adt = ops.act_dtype(precision)
    dims = fusion.dims
    lins = fusion.linears()
    m = d_emb.shape[0]
    grads: list[torch.Tensor | None] = [None] * len(lins)
    dpre_t = None  # K-major transposed dpre of the current layer
    dpre = None  # row-major dpre of the current layer (dgrad operand)
    for i in reversed(range(len(lins))):
        n_out, n_in = dims[i + 1], dims[i]
        n_out_pad, n_in_pad = _round64(n_out), _round64(n_in)
        if i == len(lins) - 1:
            dpre_t, kpad = ops.transpose_mask(d_emb, m, n_out, adt, mask=pre_last)
            if i > 0:
                dpre = ops.mask_cast_rows(d_emb, pre_last, adt)
        else:
            dpre_t, kpad = ops.transpose_mask(dpre, m, n_out_pad, adt)
        h_prev_t, _ = ops.transpose_mask(acts[i], m, n_in_pad, adt)
        rows_w = n_out if i == len(lins) - 1 else n_out_pad
        gw = torch.empty(rows_w, n_in_pad, dtype=torch.float32, device=d_emb.device)
        ops.gemm([(dpre_t, h_prev_t, kpad)], rows_w, n_in_pad, gw, DT_F32, precision=precision)
        grads[i] = gw[:n_out, :n_in].contiguous()
        if i > 0:
            # dpre_{i-1} = (dpre_i @ W_i) * relu'(h_{i-1}); W_i^T packed K-major on the fly (it is being trained)
            k_dim = n_out if i == len(lins) - 1 else n_out_pad
            w_t = torch.zeros(n_in_pad, _round64(k_dim), dtype=torch.float32, device=d_emb.device)
            w_t[:n_in, :n_out] = lins[i].weight.detach().float().t()
            w_t_packed = ops.cast_rows(w_t, adt)
            if _round64(k_dim) != k_dim:
                raise TsfmxError("fusion backward: output width must be a multiple of 64")
            nxt = ops.alloc(m, n_in_pad, adt, d_emb.device)
            ops.gemm([(dpre, w_t_packed, k_dim)], m, n_in_pad, nxt, adt, precision=precision, act=ACT_RELU_GRAD,
                     aux=_hi_view(acts[i], n_in_pad), split_off=n_in_pad)
            dpre = nxt
    return grads


def _hi_view(t: torch.Tensor, cols: int) -> torch.Tensor:
    """relu'(h) only needs the sign: the hi half of a split activation is enough."""
    return t[:, :cols] if t.shape[1] != cols else t


def fusion_forward_with_grad(fusion, ts_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
    """Stand-alone differentiable fusion (used when the fusion module is called outside MultimodalDecoder)."""
    return _FusionOnly.apply(fusion, ts_embeddings, text_embeddings, *[l.weight for l in fusion.linears()])


class _FusionOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fusion, ts, text, *weights):
        precision = PRECISIONS[fusion.precision]
        lead, d = ts.shape[:-1], ts.shape[-1]
        out, saved = fusion_forward_saving(fusion, ts.reshape(-1, d), text.reshape(-1, text.shape[-1]), precision)
        ctx.fusion, ctx.saved_acts, ctx.precision = fusion, saved, precision
        ctx.ts_requires_grad = ts.requires_grad
        return out.view(*lead, d)

    @staticmethod
    def backward(ctx, grad_out):
        g = grad_out.reshape(-1, grad_out.shape[-1]).contiguous().float()
        grads = fusion_backward(ctx.fusion, ctx.saved_acts, g, ctx.precision)
        return (None, grad_out if ctx.ts_requires_grad else None, None, *grads)

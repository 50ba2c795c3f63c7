// Whole-stack entry point: the L decoder layers of TimesFM 2.5 in ONE call over a packed weight table
// (SURVEY.md section 8(b) item 6 `timesfm_stack_fwd`, item 12 `workspace_bytes`).
//
// Replaces the loop `for layer in stacked_xf: x = layer(x, masks[..., -1], None)` of the reference adapter
// (reference tsfmx/tsfm/timesfm.py:95-98).  The per-kernel entry points stay (the training path and the decode
// steps interleave other work between them); this one exists because a binding that crosses the FFI once per kernel
// pays ~36 us of host time per launch (~13 ms for the 350 launches of a 50-layer forward).  It launches exactly the
// kernels the per-kernel path launches, in the same order, on the caller's stream, into a caller-owned workspace.
#include "common.cuh"

namespace tsfmx {
namespace {

struct Layout {
  size_t xn, qkv, attn, a, h, total;
};

inline size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// element sizes: GEMM operands (bf16, or split = two bf16) and GEMM -> norm / attention intermediates (bf16 or fp32)
inline Layout layout_for(const tsfmx_timesfm_stack* s, int64_t rows) {
  const bool fast = s->precision == TSFMX_PREC_BF16;
  const size_t op = fast ? 2 : 4, mid = fast ? 2 : 4;
  const size_t d = static_cast<size_t>(s->model_dims), ff = static_cast<size_t>(s->ff_dims), r = static_cast<size_t>(rows);
  Layout l;
  size_t off = 0;
  l.xn = off, off = align_up(off + r * d * op);
  l.qkv = off, off = align_up(off + r * 3 * d * mid);
  l.attn = off, off = align_up(off + r * d * op);
  l.a = off, off = align_up(off + r * d * mid);
  l.h = off, off = align_up(off + r * ff * op);
  l.total = off;
  return l;
}

int check_table(const tsfmx_timesfm_stack* s) {
  TSFMX_REQUIRE(s != nullptr, "timesfm_stack: NULL table");
  TSFMX_REQUIRE(s->num_layers >= 0 && (s->num_layers == 0 || s->layers != nullptr), "timesfm_stack: bad layer table");
  TSFMX_REQUIRE(s->precision == TSFMX_PREC_BF16 || s->precision == TSFMX_PREC_BF16X3, "timesfm_stack: bad precision");
  TSFMX_REQUIRE(s->model_dims > 0 && s->model_dims % 64 == 0 && s->ff_dims > 0 && s->ff_dims % 64 == 0 &&
                    s->num_heads * s->head_dim == s->model_dims,
                "timesfm_stack: model_dims %d / ff_dims %d must be multiples of 64 and heads x head_dim = model_dims",
                s->model_dims, s->ff_dims);
  return TSFMX_OK;
}

int gemm1(const void* a, int64_t lda, const void* b, int64_t ldb, int k, int64_t m, int n, int precision, int act,
          void* d, int d_dtype, cudaStream_t stream) {
  tsfmx_gemm_args g = {};
  g.m = m, g.n = n, g.num_segments = 1;
  g.seg[0].a = a, g.seg[0].lda = lda, g.seg[0].b = b, g.seg[0].ldb = ldb, g.seg[0].k = k;
  g.precision = precision, g.act = act;
  g.d = d, g.ldd = d_dtype == TSFMX_DT_BF16_SPLIT ? 2 * static_cast<int64_t>(n) : n;
  g.d_dtype = d_dtype;
  return tsfmx_gemm(&g, stream);
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" size_t tsfmx_timesfm_stack_workspace_bytes(const tsfmx_timesfm_stack* stack, int64_t batch,
                                                      int32_t num_patches) {
  if (check_table(stack) != TSFMX_OK || batch < 0 || num_patches <= 0) return 0;
  return layout_for(stack, batch * num_patches).total;
}

extern "C" int tsfmx_timesfm_stack_fwd(const tsfmx_timesfm_stack* stack, int64_t batch, int32_t num_patches,
                                       const float* x, const uint8_t* patch_mask, const int32_t* num_masked,
                                       void* workspace, size_t workspace_bytes, float* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rc = check_table(stack);
  if (rc != TSFMX_OK) return rc;
  TSFMX_REQUIRE(x != nullptr && y != nullptr && x != y, "timesfm_stack_fwd: x and y must be distinct non-NULL buffers");
  TSFMX_REQUIRE(batch >= 0 && num_patches > 0, "timesfm_stack_fwd: bad sizes");
  if (batch == 0) return TSFMX_OK;
  const int64_t rows = batch * num_patches;
  const Layout lay = layout_for(stack, rows);
  TSFMX_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total && reinterpret_cast<uintptr_t>(workspace) % 256 == 0,
                "timesfm_stack_fwd: workspace of %zu bytes (256-byte aligned) needed, got %zu", lay.total, workspace_bytes);
  const int d = stack->model_dims, ff = stack->ff_dims, prec = stack->precision;
  const bool fast = prec == TSFMX_PREC_BF16;
  const int adt = fast ? TSFMX_DT_BF16 : TSFMX_DT_BF16_SPLIT;  // GEMM operands
  const int mid = fast ? TSFMX_DT_BF16 : TSFMX_DT_F32;         // GEMM outputs read by the norm / attention kernels
  const int64_t ld_op = fast ? 1 : 2;                          // leading-dimension factor of a split operand
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void *xn = ws + lay.xn, *qkv = ws + lay.qkv, *attn = ws + lay.attn, *a = ws + lay.a, *h = ws + lay.h;
  if (stack->num_layers == 0) {
    const cudaError_t e = cudaMemcpyAsync(y, x, static_cast<size_t>(rows) * d * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) {
      set_error("timesfm_stack_fwd: %s", cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
    return TSFMX_OK;
  }
  const tsfmx_timesfm_layer* L = stack->layers;
#define TSFMX_TRY(call)            \
  do {                             \
    const int rc_ = (call);        \
    if (rc_ != TSFMX_OK) return rc_; \
  } while (0)
  TSFMX_TRY(tsfmx_rmsnorm(x, rows, d, L[0].pre_attn_ln, stack->eps, adt, xn, stream));
  for (int i = 0; i < stack->num_layers; ++i) {
    const tsfmx_timesfm_layer& w = L[i];
    const float* next_ln = i + 1 < stack->num_layers ? L[i + 1].pre_attn_ln : nullptr;
    const bool last = i + 1 == stack->num_layers;
    TSFMX_TRY(gemm1(xn, ld_op * d, w.qkv, ld_op * d, d, rows, 3 * d, prec, TSFMX_ACT_NONE, qkv, mid, stream));
    TSFMX_TRY(tsfmx_timesfm_attention(qkv, mid, batch, num_patches, stack->num_heads, stack->head_dim, patch_mask,
                                      num_masked, stack->inv_freq, w.q_ln, w.k_ln, w.q_scale, stack->eps, adt, attn,
                                      stream));
    TSFMX_TRY(gemm1(attn, ld_op * d, w.out, ld_op * d, d, rows, d, prec, TSFMX_ACT_NONE, a, mid, stream));
    TSFMX_TRY(tsfmx_norm_residual_norm(a, mid, i == 0 ? x : y, rows, d, w.post_attn_ln, w.pre_ff_ln, stack->eps, y, adt,
                                       xn, stream));
    TSFMX_TRY(gemm1(xn, ld_op * d, w.ff0, ld_op * d, d, rows, ff, prec, TSFMX_ACT_SILU, h, adt, stream));
    TSFMX_TRY(gemm1(h, ld_op * ff, w.ff1, ld_op * ff, ff, rows, d, prec, TSFMX_ACT_NONE, a, mid, stream));
    TSFMX_TRY(tsfmx_norm_residual_norm(a, mid, y, rows, d, w.post_ff_ln, next_ln, stack->eps, y, adt, last ? nullptr : xn,
                                       stream));
  }
#undef TSFMX_TRY
  return TSFMX_OK;
}

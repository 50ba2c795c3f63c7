/*
 * tsfmx_b200 — C ABI of the B200-native multimodal forecast hot path.
 *
 * Drop-in boundary for the device stages behind the TSFMx adapter API
 * (reference: /root/reference/src/tsfmx, package `tsfmx` 1.0.1).  Every entry
 * point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream
 * (passed as `void*`, i.e. a `cudaStream_t`), returns an int status
 * (0 = TSFMX_OK) and never throws.  `tsfmx_last_error()` returns the message
 * of the last failure on the calling thread.
 *
 * Ownership: the caller (PyTorch on the host side) allocates and owns every
 * buffer; the library borrows pointers for the duration of the call, enqueues
 * work on the caller's stream, performs no cudaMalloc and no host sync and
 * keeps no reference after return.  There is no CPU fallback: on a machine
 * without an sm_100 device every compute entry point fails with
 * TSFMX_ERR_NO_DEVICE.
 *
 * Tensor conventions: row-major contiguous unless a leading dimension is
 * given; masks are one byte per element, non-zero = padded (the tsfmx
 * convention, reference base.py:17).
 *
 * Activation storage ("precision" argument of the dense entry points):
 *   TSFMX_PREC_BF16    operands are bf16, accumulation fp32 (throughput mode)
 *   TSFMX_PREC_BF16X3  every operand is kept as a (hi, lo) pair of bf16 with
 *                      x ~= hi + lo (16 mantissa bits) and each product is
 *                      evaluated as hi*hi + hi*lo + lo*hi on the same tcgen05
 *                      kernel, fp32 accumulate (parity mode, <= 1e-3 relative
 *                      against the fp32 reference).
 *   A "split" matrix of logical shape [R, K] is stored as [R, 2K] bf16:
 *   columns [0, K) hold hi, columns [K, 2K) hold lo.
 */
#ifndef TSFMX_B200_H_
#define TSFMX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSFMX_ABI_VERSION 2

enum {
  TSFMX_OK = 0,
  TSFMX_ERR_INVALID_ARGUMENT = 1,
  TSFMX_ERR_CUDA = 2,
  TSFMX_ERR_NO_DEVICE = 3,
  TSFMX_ERR_UNSUPPORTED = 4
};

enum { TSFMX_PREC_BF16 = 0, TSFMX_PREC_BF16X3 = 1 };

/* element storage of an output / input matrix */
enum { TSFMX_DT_F32 = 0, TSFMX_DT_BF16 = 1, TSFMX_DT_BF16_SPLIT = 2 };

/* *_GRAD: the epilogue multiplies the accumulator (an upstream gradient) by act'(aux) — backward pass */
enum { TSFMX_ACT_NONE = 0, TSFMX_ACT_SILU = 1, TSFMX_ACT_RELU = 2, TSFMX_ACT_SILU_GRAD = 3, TSFMX_ACT_RELU_GRAD = 4 };

int tsfmx_abi_version(void);
/* sizeof(tsfmx_gemm_args) as compiled: lets a binding verify its struct layout */
int tsfmx_sizeof_gemm_args(void);
const char* tsfmx_last_error(void);
/* number of kernels this library has launched on the calling process (all threads) */
uint64_t tsfmx_launch_count(void);
/* 0 when an sm_100 device is usable for `device` (<0: current device) */
int tsfmx_device_check(int device);

/* ------------------------------------------------------------------------
 * HBM-bound stages
 * ---------------------------------------------------------------------- */

/*
 * TimesFM 2.5 patchify + running RevIN statistics + normalise + mask-concat.
 * Replaces the first half of TimesFM2p5Adapter.preprocess
 * (reference tsfmx/tsfm/timesfm.py:53-73: reshape to patches, the N-step
 * update_running_stats loop, revin, where(mask, 0, x), cat([x, mask])).
 *
 *   x          [B, C] fp32           context values
 *   mask       [B, C] u8             non-zero = padded
 *   patch_len  P (32 for TimesFM 2.5); C % P == 0, N = C / P
 *   tokens     [B*N, 2P] of tokens_dtype (F32 / BF16 / BF16_SPLIT -> [B*N, 4P])
 *              = concat(normalised values with padded->0, mask as 0/1)
 *   mu, sigma  [B, N] fp32           cumulative mean / std after each patch
 *   patch_mask [B, N] u8             mask of the LAST element of each patch
 *                                    (reference timesfm.py:97 `masks[..., -1]`)
 *   num_masked [B] int32             sum of patch_mask per series
 * Any of mu / sigma / patch_mask / num_masked may be NULL.
 */
int tsfmx_timesfm_patchify_norm(const float* x, const uint8_t* mask, int64_t batch, int32_t context,
                                int32_t patch_len, int32_t tokens_dtype, void* tokens, float* mu, float* sigma,
                                uint8_t* patch_mask, int32_t* num_masked, void* stream);

/*
 * Chronos-2 instance-norm + arcsinh + patch(16) + time-encoding concat.
 * Replaces Chronos2Model._prepare_patched_context as called by
 * Chronos2Adapter.preprocess (reference tsfmx/tsfm/chronos.py:48-52).
 *
 *   x, mask     [B, C] fp32 / u8 (non-zero = padded; tsfmx convention)
 *   patch       16; C is left-padded (NaN) to Cp = ceil(C/patch)*patch, N = Cp/patch
 *   patched     [B*N, out_cols] of out_dtype = [time_enc | values | mask | zero fill]; out_cols >= 3*patch
 *               (64 for the 48-wide Chronos-2 patch so the row is one GEMM k-block)
 *   attn_mask   [B, N] u8   1 = patch has at least one observed point
 *   loc, scale  [B] fp32
 */
int tsfmx_chronos2_patchify_norm(const float* x, const uint8_t* mask, int64_t batch, int32_t context,
                                 int32_t patch, int32_t use_arcsinh, float time_encoding_scale,
                                 int32_t out_dtype, int32_t out_cols, void* patched, uint8_t* attn_mask,
                                 float* loc, float* scale, void* stream);

/*
 * Chronos-T5 MeanScaleUniformBins tokeniser (north-star item; upstream
 * chronos.MeanScaleUniformBins.context_input_transform, not in the reference).
 *
 *   x         [B, C] fp32, NaN = missing
 *   ids       [B, C+1] int64   token ids, EOS (1) appended, PAD (0) for NaN
 *   attn_mask [B, C+1] u8
 *   scale     [B] fp32          mean(|x|) over observed points (1 if not > 0)
 *   boundaries[n_boundaries] fp32 device pointer, ascending (bucketize right=True)
 */
int tsfmx_chronos_t5_tokenize(const float* x, int64_t batch, int32_t context, const float* boundaries,
                              int32_t n_boundaries, int32_t n_special, int32_t n_tokens, int32_t pad_id,
                              int32_t eos_id, int64_t* ids, uint8_t* attn_mask, float* scale, void* stream);

/* ids [B, L] int64 -> values [B, L] fp32 = centers[clamp(id - n_special - 1, 0, n_centers-1)] * scale[b] */
int tsfmx_chronos_t5_dequantize(const int64_t* ids, int64_t batch, int32_t length, const float* centers,
                                int32_t n_centers, int32_t n_special, const float* scale, float* values,
                                void* stream);

/* fp32 [rows, cols] (leading dim ld_in) -> bf16 or split-bf16 [rows, cols | 2*cols] */
int tsfmx_cast_rows(const float* in, int64_t rows, int32_t cols, int64_t ld_in, int32_t out_dtype, void* out,
                    void* stream);

/* ------------------------------------------------------------------------
 * Dense stages (tcgen05 / TMEM GEMM fed by TMA)
 * ---------------------------------------------------------------------- */

/* One K-segment of a GEMM: contributes A_s[M, k] * B_s[N, k]^T to the accumulator. */
typedef struct {
  const void* a;  /* bf16 [M, k] (or split [M, 2k]); K-major */
  int64_t lda;    /* elements */
  const void* b;  /* bf16 [N, k] (or split [N, 2k]); K-major, e.g. nn.Linear.weight */
  int64_t ldb;    /* elements */
  int32_t k;      /* logical K of this segment; multiple of 64 */
  int32_t reserved;
} tsfmx_gemm_segment;

/*
 * D = epilogue( sum_s A_s * B_s^T )   with epilogue, in this order:
 *   v = acc + bias[col]; [pre_act[row, col] = v;] v = act(v)  (or v *= act'(aux[row, col]) for *_GRAD);
 *   v = v * row_scale[row] + row_shift[row]; v += residual[row, col]; store columns < n_store as d_dtype.
 * Used for every Linear on the path: TimesFM tokenizer / transformer / head
 * ResidualBlocks (two segments: hidden path + residual path), MultimodalFusion
 * (reference fusion.py:46-47: relu epilogue + residual add), Chronos-2 blocks.
 */
typedef struct {
  int64_t m;
  int32_t n;
  int32_t num_segments; /* 1 or 2 */
  tsfmx_gemm_segment seg[2];
  int32_t precision;      /* TSFMX_PREC_* : whether A/B are split matrices */
  int32_t act;            /* TSFMX_ACT_* */
  const float* bias;      /* [n] or NULL */
  const float* row_scale; /* [m] or NULL */
  const float* row_shift; /* [m] or NULL */
  const float* residual;  /* fp32 [m, n] (ld = ldr) or NULL */
  int64_t ldr;
  void* d;         /* output */
  int64_t ldd;     /* elements of d_dtype (for BF16_SPLIT: lo half starts at column split_off) */
  int32_t d_dtype; /* TSFMX_DT_* */
  int32_t n_store; /* store only columns < n_store (0 = n) */
  int32_t split_off; /* BF16_SPLIT: column offset of the lo half (0 = n) */
  int32_t aux_dtype;      /* TSFMX_DT_F32 or TSFMX_DT_BF16 */
  const void* aux;        /* [m, >= n_store]: SILU_GRAD -> saved pre-activation u; RELU_GRAD -> pre- or post-activation */
  int64_t ld_aux;
  void* pre_act;          /* optional [m, >= n_store]: acc + bias before the activation (training forward) */
  int64_t ld_pre;
  int32_t pre_act_dtype;  /* TSFMX_DT_F32 or TSFMX_DT_BF16 */
  int32_t reserved;
} tsfmx_gemm_args;

int tsfmx_gemm(const tsfmx_gemm_args* args, void* stream);
/* test / tuning hook: 0 = automatic, 1 = one CTA per tile (UMMA 128xN), 2 = CTA pairs (cta_group::2, UMMA 256xN) */
int tsfmx_gemm_set_cta_group(int cta_group);
/* test / tuning hook for split-K (used automatically for plain fp32 outputs with fewer tiles than SMs, i.e. the weight
 * gradients dW = dY^T X of the full fine-tuning path, reference trainer.py:78-79, where K = tokens): 0 = automatic,
 * 1 = never, 2..16 = that many splits wherever splitting is legal */
int tsfmx_gemm_set_split_k(int mode);

/*
 * Weight gradient of y = x W^T (full fine-tuning, reference trainer.py:78-79,123: every Linear of the adapter trains):
 *   dw [n_out, k_in] fp32 = dy[rows, n_out]^T x[rows, k_in],  dy / x bf16 row-major as the backward pass left them.
 * K = tokens: the operands are MN-major for this product and are consumed as such (TMA boxes of 64 features x 64
 * tokens, MN-major UMMA descriptors) - no transposed copies; tokens beyond `rows` are zero-filled by TMA; split-K as
 * in tsfmx_gemm.  ld_dy / ld_x in elements (multiples of 8), pointers 16-byte aligned, n_out / k_in multiples of 8.
 */
int tsfmx_gemm_wgrad(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t rows, int32_t n_out,
                     int32_t k_in, float* dw, int64_t ld_dw, void* stream);

/*
 * GEMM with the norm / residual junction of a transformer layer fused into its epilogue (one launch instead
 * of tsfmx_gemm + tsfmx_norm_residual_norm; the GEMM result never goes to HBM):
 *   a = A W^T;  y = RMSNorm(a) * w_post + x  (w_post NULL: y = a + x);  yn = RMSNorm(y) * w_next  (w_next NULL: yn = y)
 * Replaces attn.out / ff1 + post_ln + residual + next pre_ln of upstream timesfm Transformer.forward (called at
 * reference tsfmx/tsfm/timesfm.py:97; HF twin modeling_timesfm2_5.py:378-388).  n / 256 CTAs form a cluster that owns
 * a full 128-row x n panel; row statistics are exchanged through distributed shared memory; x is read and y written
 * through TMA-staged 128 x 32 chunks.  Opt-in (the adapters' fused_norm flag): measured at par with tsfmx_gemm +
 * tsfmx_norm_residual_norm, not faster.
 *   seg: one K-segment (split operands when precision = BF16X3); n in {512, 768, 1024, 1280}
 *   x, y fp32 [m, n] with contiguous rows of n floats, 16-byte aligned (y may alias x); yn [m, n] of yn_dtype or NULL.
 */
int tsfmx_gemm_rownorm(const tsfmx_gemm_segment* seg, int64_t m, int32_t n, int32_t precision, const float* w_post,
                       const float* w_next, const float* x, float* y, int32_t yn_dtype, void* yn, float eps,
                       void* stream);

/* y = x * rsqrt(mean(x^2) + eps) * w, rows of `cols` fp32 -> bf16 / split / f32. */
int tsfmx_rmsnorm(const float* x, int64_t rows, int32_t cols, const float* w, float eps, int32_t out_dtype,
                  void* out, void* stream);

/*
 * Fused post-norm + residual + next pre-norm of a TimesFM 2.5 layer
 * (upstream Transformer.forward: `post_ln(a) + x` followed by the next
 * `pre_ln`; HF twin modeling_timesfm2_5.py:378-388):
 *   y   = rmsnorm(a) * w_post + x          (fp32, written to y; may alias x)
 *   yn  = rmsnorm(y) * w_next              (bf16 / split; w_next NULL -> yn = y cast)
 *   a is a_dtype (F32 or BF16) [rows, cols].
 */
int tsfmx_norm_residual_norm(const void* a, int32_t a_dtype, const float* x, int64_t rows, int32_t cols,
                             const float* w_post, const float* w_next, float eps, float* y, int32_t yn_dtype,
                             void* yn, void* stream);

/*
 * TimesFM 2.5 attention core for one layer (upstream MultiHeadAttention after
 * qkv_proj; HF twin modeling_timesfm2_5.py:304-346): RoPE(q, k) with
 * position = n - num_masked[b], RMSNorm over head_dim on q and k, per-dim
 * softplus query scale, softmax(q k^T + causal & key-not-padded mask) v.
 *
 *   qkv        [B*N, 3*H*hd] of qkv_dtype (F32 or BF16): [q | k | v], head-major inside
 *   patch_mask [B, N] u8, num_masked [B] int32
 *   inv_freq   [hd/2] fp32;  q_ln_w, k_ln_w [hd];  q_scale [hd] = softplus(per_dim_scale)*1.442695041/sqrt(hd)
 *   out        [B*N, H*hd] of out_dtype (BF16 / BF16_SPLIT / F32)
 * A query whose keys are all masked attends uniformly to all N keys (what the
 * additive finfo.min mask of the reference produces).
 */
int tsfmx_timesfm_attention(const void* qkv, int32_t qkv_dtype, int64_t batch, int32_t num_patches,
                            int32_t num_heads, int32_t head_dim, const uint8_t* patch_mask,
                            const int32_t* num_masked, const float* inv_freq, const float* q_ln_w,
                            const float* k_ln_w, const float* q_scale, float eps, int32_t out_dtype, void* out,
                            void* stream);

/*
 * Chronos-2 time self-attention core of one encoder block (upstream Chronos2Model.encoder, driven by
 * Chronos2Adapter.forward, reference tsfmx/tsfm/chronos.py:119-123): RoPE(q, k) with positions arange(T),
 * no 1/sqrt(d) scaling, bidirectional, additive key mask, fp32 softmax.
 *   qkv [B*T, 3*H*hd] f32 / bf16 = [q | k | v];  key_mask [B, T] u8, non-zero = attendable;  inv_freq [hd/2]
 *   out [B*T, H*hd] of out_dtype.  A row whose keys are all masked gets uniform weights (finfo.min semantics).
 * The block's group self-attention needs no kernel: with group_ids = arange(B) (chronos.py:117) it equals
 * h + (W_o W_v) RMSNorm(h), one tsfmx_gemm.
 */
int tsfmx_encoder_attention(const void* qkv, int32_t qkv_dtype, int64_t batch, int32_t seq, int32_t num_heads,
                            int32_t head_dim, const uint8_t* key_mask, const float* inv_freq, int32_t out_dtype,
                            void* out, void* stream);

/*
 * Tensor-core variant of tsfmx_encoder_attention for the throughput mode: qkv and out bf16, head_dim 64, seq <= 208
 * (Chronos-2: T = ctx/16 + 65 = 97 at ctx 512, 193 at ctx 2048).  Same semantics (positions arange(T), no 1/sqrt(d),
 * bidirectional, key mask, all-masked rows -> uniform weights); rope_table [seq, head_dim/2, 2] fp32 = (cos, sin) of
 * position * inv_freq, filled once by tsfmx_rope_table and reused by every block and every call.
 */
int tsfmx_rope_table(const float* inv_freq, int32_t half_dim, int32_t seq, float* table, void* stream);
int tsfmx_encoder_attention_mma(const void* qkv, int64_t batch, int32_t seq, int32_t num_heads, int32_t head_dim,
                                const uint8_t* key_mask, const float* rope_table, void* out, void* stream);

/*
 * Column reductions of the full fine-tune ("baseline" mode, reference tsfmx/trainer.py:78-79,123): normalize != 0:
 * out[c] += sum_r g[r, c] * v[r, c] * rsqrt(mean(v[r]^2) + eps) (gradient of an RMSNorm scale); normalize == 0:
 * out[c] += sum_r g[r, c] (gradient of a bias).  v, g [rows, cols] f32 / bf16, cols in {1280, 768}; out fp32 [cols],
 * accumulated into (zero it first).
 */
int tsfmx_colsum_wgrad(const void* v, int32_t v_dtype, const void* g, int32_t g_dtype, int64_t rows, int32_t cols, float eps,
                       int32_t normalize, float* out, void* stream);

/*
 * Backward of tsfmx_encoder_attention: d_out [B*T, H*64] -> dqkv [B*T, 3*H*64] = [dq | dk | dv] (fusion fine-tune
 * through the frozen Chronos-2 encoder; the reference trains the fusion module with either adapter,
 * scripts/tune_time_mmd_sweep.py:124-126).  fp32 arithmetic; stats_workspace: B*H*T float4 scratch (row max, 1 / sum,
 * delta, has-key flag), 16-byte aligned.  All-masked series pass gradient like the reference's additive mask.
 */
int tsfmx_encoder_attention_bwd(const void* qkv, int32_t qkv_dtype, const void* d_out, int32_t dout_dtype, int64_t batch,
                                int32_t seq, int32_t num_heads, int32_t head_dim, const uint8_t* key_mask,
                                const float* rope_table, float* stats_workspace, int32_t dqkv_dtype, void* dqkv,
                                void* stream);

/*
 * Chronos-T5 backbone stages (BASELINE.json configs[2]; upstream chronos.ChronosModel -> transformers T5ForConditionalGeneration,
 * HF twin transformers/models/t5/modeling_t5.py; not part of the reference, which only wraps Chronos-2).
 *
 * tsfmx_embed_rows: out[r, :] = table[ids[r], :] (fp32 table [vocab, dims], dims % 4 == 0).
 *
 * tsfmx_t5_attention: general T5 attention core, fp32 SIMT (parity mode, single-query decoding over the KV cache,
 *   teacher forcing).  For series b, query row i (position q_pos0 + i) and key j (position j):
 *     score = q . k  (no 1/sqrt(d))  + bias[h, (j - q_pos) + bias_zero]   if bias != NULL (index clamped to the table)
 *     key j takes part iff (key_mask == NULL or key_mask[b, j] != 0) and (not causal or j <= q_pos);
 *     a query whose admissible keys are all masked attends uniformly to them (finfo.min semantics).
 *   q [B, tq] rows of stride ldq (series stride q_batch_stride, both in elements), k / v likewise with tk rows;
 *   head h owns columns [64 h, 64 h + 64) of every row; out rows of stride ldo, width num_heads * 64.
 *   kv_batch_div >= 1: query series b reads the keys / values / mask of series b / kv_batch_div (the sample paths of
 *   one series share the encoder-side keys and values).
 *
 * tsfmx_t5_encoder_attention_mma: throughput mode of the encoder self-attention (bidirectional, key mask, bias table
 *   [num_heads, 2 seq - 1] indexed by (key - query) + seq - 1): qkv bf16 [B*seq, 3*H*64] = [q | k | v], out bf16
 *   [B*seq, H*64]; K and V of one (series, head) stay in shared memory, online softmax over 64-key chunks; seq <= 704.
 */
/*
 * Next-token choice of a Chronos-T5 decode step in one kernel (upstream ChronosModel.forward -> generate(do_sample=True,
 * top_k=50, temperature=1.0), HF generation/logits_process.py TopKLogitsWarper + multinomial): logits [rows, vocab] fp32
 * (not modified), `banned_id` (EOS while min_new_tokens holds; < 0: none) excluded, logits / temperature, the top_k
 * largest kept (<= 0: all), softmax, inverse-CDF sampling in id order with uniform[row] in [0, 1).  top_k = 1 is greedy
 * decoding (first maximum; uniform may be NULL).  out [rows] int64.
 */
int tsfmx_t5_sample_topk(const float* logits, int64_t rows, int32_t vocab, int32_t banned_id, float temperature,
                         int32_t top_k, const float* uniform, int64_t* out, void* stream);

int tsfmx_embed_rows(const int64_t* ids, int64_t rows, int32_t dims, int32_t vocab, const float* table, float* out,
                     void* stream);
int tsfmx_t5_attention(const void* q, int32_t q_dtype, int64_t ldq, int64_t q_batch_stride, const void* k, const void* v,
                       int32_t kv_dtype, int64_t ldk, int64_t ldv, int64_t kv_batch_stride, int64_t batch, int32_t tq,
                       int32_t tk, int32_t num_heads, int32_t head_dim, int32_t q_pos0, int32_t causal,
                       const uint8_t* key_mask, const float* bias, int32_t bias_len, int32_t bias_zero,
                       int32_t out_dtype, void* out, int64_t ldo, int64_t o_batch_stride, int32_t kv_batch_div,
                       void* stream);
int tsfmx_t5_encoder_attention_mma(const void* qkv, int64_t batch, int32_t seq, int32_t num_heads, int32_t head_dim,
                                   const uint8_t* key_mask, const float* bias, void* out, void* stream);

/*
 * Chronos-2 output epilogue (reference chronos.py:159-169): preds [B*np, Q*patch] (patch-major rows of the
 * output ResidualBlock, np = ceil(horizon / patch)) -> out [B, horizon, Q] = sinh(x) * scale[b] + loc[b].
 */
int tsfmx_chronos2_finalize(const float* preds, int64_t batch, int32_t num_patches_used, int32_t num_quantiles,
                            int32_t patch, int32_t horizon, int32_t use_arcsinh, const float* loc, const float* scale,
                            float* out, void* stream);

/* ------------------------------------------------------------------------
 * Backward pass of the fusion fine-tune step (reference tsfmx/trainer.py:200-219; the adapter is frozen,
 * trainer.py:76-77, so only activation gradients flow through the backbone).  The dgrad GEMMs are
 * tsfmx_gemm on pre-transposed weights with the *_GRAD epilogues.
 * ---------------------------------------------------------------------- */

/*
 * One norm / residual junction of a TimesFM 2.5 layer, backwards:
 *   g_total = g_res + RMSNorm_bwd(v1, w1, g1)      (either term may be absent: g_res NULL / v1 NULL)
 *   g2      = RMSNorm_bwd(v2, w2, g_total)         (skipped when v2 is NULL)
 * with RMSNorm_bwd(v, w, g) = r (w g) - v r^3 mean(v w g), r = (mean(v^2) + eps)^-1/2.
 * v1 / g1 / v2 are f32 or bf16 [rows, cols]; g_total fp32 (may alias g_res); g2 of g2_dtype.
 */
int tsfmx_rmsnorm_bwd_chain(const float* g_res, const void* v1, int32_t v1_dtype, const float* w1, const void* g1,
                            int32_t g1_dtype, const void* v2, int32_t v2_dtype, const float* w2, int64_t rows,
                            int32_t cols, float eps, float* g_total, int32_t g2_dtype, void* g2, void* stream);
/* The same pass with the gradients of the two norm scales riding along (full fine-tuning, reference trainer.py:78-79):
 *   dw1[c] += sum_r g1[r, c] * v1_hat[r, c],   dw2[c] += sum_r g_total[r, c] * v2_hat[r, c]
 * (v_hat = RMS-normalised v; dw1 / dw2 fp32 [cols], accumulated into, either may be NULL). */
int tsfmx_rmsnorm_bwd_chain_wgrad(const float* g_res, const void* v1, int32_t v1_dtype, const float* w1, const void* g1,
                                  int32_t g1_dtype, const void* v2, int32_t v2_dtype, const float* w2, int64_t rows,
                                  int32_t cols, float eps, float* g_total, int32_t g2_dtype, void* g2, float* dw1,
                                  float* dw2, void* stream);

/*
 * Gradient of tsfmx_timesfm_attention w.r.t. its qkv input: recomputes the conditioned q / k and the
 * probabilities, then dV = P^T dO, dS = P (dO V^T - rowsum), dq' = dS k', dk' = dS^T q' and back through
 * per-dim scale, RMSNorm and RoPE.  d_out [B*N, H*hd] f32 / bf16; dqkv [B*N, 3*H*hd] of dqkv_dtype.
  * dparams (NULL when the adapter is frozen): fp32 [2 * head_dim], accumulated into - [0, hd) gets
 * d/d(q_ln_w * q_scale) = sum dq' * qhat, [hd, 2 hd) gets d/d(k_ln_w) = sum dk' * khat (full fine-tuning, "baseline" mode).
 */
int tsfmx_timesfm_attention_bwd(const void* qkv, int32_t qkv_dtype, const void* d_out, int32_t dout_dtype,
                                int64_t batch, int32_t num_patches, int32_t num_heads, int32_t head_dim,
                                const uint8_t* patch_mask, const int32_t* num_masked, const float* inv_freq,
                                const float* q_ln_w, const float* k_ln_w, const float* q_scale, float eps,
                                int32_t dqkv_dtype, void* dqkv, float* dparams, void* stream);

/*
 * out[c, r] = (mask == NULL || mask[r, c] > 0) ? in[r, c] : 0 for r < rows, zero for rows <= r < ld_out.
 * Produces the K-major operands of the fusion weight gradient dW[n, k] = sum_tokens dpre[t, n] text[t, k]
 * (reference: autograd of fusion.py:46-47) with K = tokens padded to ld_out (a multiple of 64).
 * in / mask: f32, bf16 or split [rows, cols]; out: bf16 [cols, ld_out] or split [cols, 2*ld_out].
 */
int tsfmx_transpose_mask(const void* in, int32_t in_dtype, int64_t rows, int32_t cols, int64_t ld_in, const void* mask,
                         int32_t mask_dtype, int64_t ld_mask, int32_t out_dtype, void* out, int64_t ld_out,
                         void* stream);

/* out[r, c] = mask[r, c] > 0 ? in[r, c] : 0; fp32 in -> f32 / bf16 / split out (relu' gate of the fusion dgrad) */
int tsfmx_mask_cast_rows(const float* in, int64_t rows, int32_t cols, const void* mask, int32_t mask_dtype,
                         int64_t ld_mask, int32_t out_dtype, void* out, void* stream);

/* ------------------------------------------------------------------------
 * Whole-stack entry point (one FFI crossing per forward instead of one per kernel)
 * ---------------------------------------------------------------------- */

/* Weights of one TimesFM 2.5 decoder layer, as the library consumes them: Linear weights K-major [out, in] packed to
 * bf16 (precision BF16) or split bf16 [out, 2 * in] (BF16X3) with tsfmx_cast_rows; norm scales fp32. */
typedef struct {
  const void* qkv;   /* [3 D, D]   attn.qkv_proj.weight */
  const void* out;   /* [D, D]     attn.out.weight */
  const void* ff0;   /* [F, D]     ff0.weight */
  const void* ff1;   /* [D, F]     ff1.weight */
  const float* pre_attn_ln;  /* [D] */
  const float* post_attn_ln; /* [D] */
  const float* pre_ff_ln;    /* [D] */
  const float* post_ff_ln;   /* [D] */
  const float* q_ln;    /* [hd] attn.query_ln.scale */
  const float* k_ln;    /* [hd] attn.key_ln.scale */
  const float* q_scale; /* [hd] softplus(per_dim_scale) * 1.442695041 / sqrt(hd) */
} tsfmx_timesfm_layer;

typedef struct {
  int32_t num_layers, model_dims, num_heads, head_dim, ff_dims;
  int32_t precision; /* TSFMX_PREC_* */
  float eps;
  int32_t reserved;
  const float* inv_freq;             /* [hd / 2] device */
  const tsfmx_timesfm_layer* layers; /* HOST array of num_layers entries (device pointers inside) */
} tsfmx_timesfm_stack;

/* Bytes of scratch tsfmx_timesfm_stack_fwd needs for `batch` series of `num_patches` tokens (0 on a bad table). */
size_t tsfmx_timesfm_stack_workspace_bytes(const tsfmx_timesfm_stack* stack, int64_t batch, int32_t num_patches);

/*
 * All decoder layers of TimesFM 2.5: replaces `for layer in stacked_xf: x = layer(x, masks[..., -1], None)`
 * (reference tsfmx/tsfm/timesfm.py:95-98).  Per layer: qkv GEMM -> attention -> out GEMM -> post-norm + residual +
 * pre-norm -> ff0 GEMM (SiLU) -> ff1 GEMM -> post-norm + residual + next pre-norm; the same kernels, in the same
 * order, as the per-kernel entry points above.
 *   x [B*N, D] fp32 (input embeddings, read only), patch_mask [B, N] u8 / num_masked [B] as in tsfmx_timesfm_attention,
 *   workspace: caller-owned, 256-byte aligned, >= tsfmx_timesfm_stack_workspace_bytes(...);  y [B*N, D] fp32 (!= x).
 */
int tsfmx_timesfm_stack_fwd(const tsfmx_timesfm_stack* stack, int64_t batch, int32_t num_patches, const float* x,
                            const uint8_t* patch_mask, const int32_t* num_masked, void* workspace,
                            size_t workspace_bytes, float* y, void* stream);

/* ------------------------------------------------------------------------
 * TimesFM 2.5 autoregressive decode (horizon > 128) and forecast extras.
 * Beyond the reference adapter, which raises for horizon > output_patch_len
 * (reference tsfmx/tsfm/timesfm.py:116-119); follows upstream timesfm's decode loop
 * (timesfm @ 8a755c9, TimesFM_2p5_200M_torch_module.decode) and the HF port's
 * forecast extras (transformers modeling_timesfm2_5.py:797-837).
 * ---------------------------------------------------------------------- */
#define TSFMX_MAX_KV_REGIONS 16

/*
 * The `patches` new input patches of a decode step (the previous 128-step point forecast), with the running
 * RevIN statistics CONTINUED from the context: state_n / state_mu / state_sigma [B] fp32 are read and updated
 * in place (update_running_stats with an all-valid patch).  Element (b, t) of the new values is read from
 * x[b * x_series_stride + t * x_elem_stride] (a channel of a [B, steps, Q] forecast needs no gather).
 *   tokens [B * patches, 2 * patch_len] of tokens_dtype ([normalised values | mask = 0]); mu, sigma [B, patches].
 */
int tsfmx_timesfm_patchify_continue(const float* x, int64_t x_series_stride, int64_t x_elem_stride, int64_t batch,
                                    int32_t patches, int32_t patch_len, float* state_n, float* state_mu,
                                    float* state_sigma, int32_t tokens_dtype, void* tokens, float* mu, float* sigma,
                                    void* stream);

/*
 * Attention of a decode step: the tokens of the LAST region are the queries, every token of every region a key
 * (causal).  Region r is a raw qkv matrix [B * region_tokens[r], 3 * H * hd] (f32 or bf16) exactly as
 * tsfmx_gemm left it - region 0 the prefill's, then one per earlier decode step - so the "KV cache" needs no
 * copy.  region_ptrs / region_tokens are HOST arrays (num_regions <= TSFMX_MAX_KV_REGIONS).  patch_mask
 * [B, n_ctx] / num_masked [B] describe the left padding of region 0 as in tsfmx_timesfm_attention.  rope_table
 * [rope_len, hd / 2, 2] fp32 = (cos, sin)(position * inv_freq) from tsfmx_rope_table (positions beyond it, or
 * every |position|: rope_len >= max(total tokens, n_ctx).
 *   out [B * q_tokens, H * hd] of out_dtype.
 */
int tsfmx_timesfm_attention_decode(const void* const* region_ptrs, const int32_t* region_tokens, int32_t num_regions,
                                   int32_t qkv_dtype, int64_t batch, int32_t num_heads, int32_t head_dim, int32_t n_ctx,
                                   const uint8_t* patch_mask, const int32_t* num_masked, const float* rope_table,
                                   int32_t rope_len, const float* inv_freq, const float* q_ln_w, const float* k_ln_w,
                                   const float* q_scale, float eps, int32_t out_dtype, void* out, void* stream);

/*
 * Forecast extras in one pass (HF modeling_timesfm2_5.py:797-837): flip-invariance combination
 * (x - flip_quantiles(x_neg)) / 2 of the point / quantile forecasts and of the quantile spread, continuous
 * quantile head (channel c != 0, decode_index takes spread[c] - spread[decode_index] + pf[decode_index]),
 * horizon slice, clamp at zero for series whose context is non-negative.
 *   pf     [(1 + flip) * B, pf_steps, Q] fp32      rows B.. are the forecasts of the negated inputs
 *   spread [(1 + flip) * B, spread_steps, Q] fp32  or NULL;   inputs [B, context] fp32 (or NULL)
 *   out    [B, horizon, Q] fp32
 */
int tsfmx_timesfm_forecast_finalize(const float* pf, const float* spread, const float* inputs, int64_t batch,
                                    int32_t context, int32_t pf_steps, int32_t spread_steps, int32_t num_outputs,
                                    int32_t horizon, int32_t decode_index, int32_t flip,
                                    int32_t use_continuous_quantile_head, int32_t infer_is_positive, float* out,
                                    void* stream);

/* tuning hook (A/B runs): key 0 = series per warp tile of timesfm_patchify_norm, key 1 = its warps per block,
 * key 2 / key 3 != 0 force the generic fallback kernel of chronos_t5_tokenize / timesfm_patchify_norm, key 5 != 0 the
 * general tsfmx_t5_attention kernel also for a decode step's cross-attention (0 = default) */
int tsfmx_tune(int32_t key, int32_t value);

/* test hook: non-zero forces the fp32 SIMT attention kernel even where the tensor-core kernel applies */
int tsfmx_attention_force_simt(int on);

#ifdef __cplusplus
}
#endif

#endif /* TSFMX_B200_H_ */

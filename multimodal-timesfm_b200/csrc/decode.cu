// TimesFM 2.5 autoregressive decode (horizon > 128) and forecast extras - SURVEY.md section 8(f) row 1.
//
// The reference adapter refuses horizon > output_patch_len (reference tsfmx/tsfm/timesfm.py:116-119); upstream timesfm
// forecasts longer horizons by feeding the previous 128-step point forecast back as 4 new input patches whose running
// RevIN statistics continue the context's, and decoding them against a KV cache.  Three kernels make that step cheap:
//
//   timesfm_patchify_continue_kernel : the 4 new patches of every series -> tokenizer input rows [values | mask = 0],
//                                      running (n, mu, sigma) continued in place (same merge formula as the prefill)
//   timesfm_attention_decode_kernel  : attention of the NEW tokens against every earlier token.  The "KV cache" is
//                                      simply the raw qkv matrices earlier launches left in HBM: the prefill's
//                                      [B * N, 3 D] and one [B * 4, 3 D] per decode step, passed as a list of regions
//                                      (no copy, no re-layout).  Keys are conditioned (RoPE, RMSNorm, k_ln) on the
//                                      fly in 32-key tiles with an online softmax, so any context length fits.
//   timesfm_forecast_finalize_kernel : flip-invariance combination, continuous quantile head, positivity clamp and
//                                      horizon slice (HF modeling_timesfm2_5.py:797-837) in one pass.
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

// ------------------------------------------------------------------------------------------------ patchify (continue)
// One warp per series; lane l holds element l of the current patch.  update_running_stats with an all-valid patch
// (HF twin modeling_timesfm2_5.py:528-568): inc_n = P.
template <int OUT>
__global__ void timesfm_patchify_continue_kernel(const float* __restrict__ x, int64_t x_series_stride,
                                                 int64_t x_elem_stride, int64_t batch, int patches, float* state_n,
                                                 float* state_mu, float* state_sigma, void* __restrict__ tokens,
                                                 float* __restrict__ mu_out, float* __restrict__ sigma_out) {
  constexpr int P = 32;
  const int lane = threadIdx.x & 31;
  const int64_t b = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (b >= batch) return;
  float n = state_n[b], mu = state_mu[b], sigma = state_sigma[b];
  for (int p = 0; p < patches; ++p) {
    const float v = x[b * x_series_stride + static_cast<int64_t>(p * P + lane) * x_elem_stride];
    const float inc_n = static_cast<float>(P);
    const float inc_mu = warp_sum(v) / inc_n;
    const float c = v - inc_mu;
    const float inc_var = warp_sum(c * c) / inc_n;
    const float inc_sigma = sqrtf(fmaxf(inc_var, 0.f));
    const float new_n = n + inc_n;
    const float new_mu = (n * mu + inc_mu * inc_n) / new_n;
    const float t1 = n * (sigma * sigma);
    const float t2 = inc_n * (inc_sigma * inc_sigma);
    const float t3 = n * ((mu - new_mu) * (mu - new_mu));
    const float t4 = inc_n * ((inc_mu - new_mu) * (inc_mu - new_mu));
    const float new_var = (((t1 + t2) + t3) + t4) / new_n;
    n = new_n, mu = new_mu, sigma = sqrtf(fmaxf(new_var, 0.f));
    const float normed = (v - mu) / (sigma < 1e-6f ? 1.0f : sigma);
    const int64_t row = b * patches + p;
    if constexpr (OUT == TSFMX_DT_F32) {
      float* t = reinterpret_cast<float*>(tokens) + row * 2 * P;
      t[lane] = normed;
      t[P + lane] = 0.f;
    } else if constexpr (OUT == TSFMX_DT_BF16) {
      __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(tokens) + row * 2 * P;
      t[lane] = __float2bfloat16_rn(normed);
      t[P + lane] = __float2bfloat16_rn(0.f);
    } else {  // split: [hi (2P) | lo (2P)]
      __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(tokens) + row * 4 * P;
      __nv_bfloat16 hi, lo;
      split_bf16(normed, hi, lo);
      t[lane] = hi;
      t[P + lane] = __float2bfloat16_rn(0.f);
      t[2 * P + lane] = lo;
      t[3 * P + lane] = __float2bfloat16_rn(0.f);
    }
    if (lane == 0) {
      mu_out[row] = mu;
      sigma_out[row] = sigma;
    }
  }
  if (lane == 0) {
    state_n[b] = n;
    state_mu[b] = mu;
    state_sigma[b] = sigma;
  }
}

// ------------------------------------------------------------------------------------------------ decode attention
constexpr int MAX_REGIONS = TSFMX_MAX_KV_REGIONS;

struct KvRegions {
  const void* ptr[MAX_REGIONS];
  int32_t tokens[MAX_REGIONS];  // tokens per series in the region
  int32_t count;
};

template <int QKV_BF16>
__device__ __forceinline__ float ld_qkv(const void* p, int64_t idx) {
  if constexpr (QKV_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  else return reinterpret_cast<const float*>(p)[idx];
}

// One warp per (series, head).  MQ = number of new tokens (queries) = tokens of the last region.
template <int HD, int MQ, int QKV_BF16, int OUT>
__global__ void timesfm_attention_decode_kernel(KvRegions regions, int64_t batch, int num_heads, int n_ctx,
                                                const uint8_t* __restrict__ patch_mask,
                                                const int32_t* __restrict__ num_masked,
                                                const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w,
                                                const float* __restrict__ k_ln_w, const float* __restrict__ q_scale,
                                                float eps, int rope_rows, void* __restrict__ out) {
  constexpr int DPL = (HD + 31) / 32;
  constexpr int HALF = HD / 2;
  constexpr int LDS = HD + 1;
  constexpr int KT = 32;  // keys per tile
  extern __shared__ float smem_dec[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PER_WARP = (MQ + 2 * KT) * LDS + MQ * KT;
  // per-block table of (cos, sin)(pos * inv_freq) for pos in [-n_ctx, total_tokens): every (series, head) of the block
  // looks its rotations up instead of evaluating 3 sincosf per lane and key row (the kernel was SFU / range-reduction
  // bound: 68 key rows x 80 rotations per pair at ctx 2048).  rope_rows = 0: table does not fit, evaluate directly.
  float2* s_rope = reinterpret_cast<float2*>(smem_dec + warps_per_block * PER_WARP);
  float* sQ = smem_dec + warp * PER_WARP;
  float* sK = sQ + MQ * LDS;
  float* sV = sK + KT * LDS;
  float* sP = sV + KT * LDS;  // [MQ][KT]
  const int width = num_heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  int total_tokens = 0;
  for (int r = 0; r < regions.count; ++r) total_tokens += regions.tokens[r];
  const int q_pos0 = total_tokens - MQ;  // sequence index of the first new token
  const void* q_region = regions.ptr[regions.count - 1];
  if (rope_rows > 0) {
    for (int i = threadIdx.x; i < rope_rows * HALF; i += blockDim.x) {
      const int p = i / HALF, f = i - p * HALF;
      float sn, cs;
      sincosf(static_cast<float>(p - n_ctx) * __ldg(inv_freq + f), &sn, &cs);
      s_rope[i] = make_float2(cs, sn);
    }
    __syncthreads();
  }

  // RoPE (rotate-half) + RMSNorm over head_dim of one staged row, in place; `scale_w` = per-dim weight
  auto condition_row = [&](float* row, int ipos, const float* w1, const float* w2) {
    float r[DPL];
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < DPL; ++t) {
      const int d = lane + 32 * t;
      r[t] = 0.f;
      if (d < HD) {
        const int f = d < HALF ? d : d - HALF;
        float sn, cs;
        if (rope_rows > 0) {
          const float2 e = s_rope[(ipos + n_ctx) * HALF + f];
          cs = e.x, sn = e.y;
        } else {
          sincosf(static_cast<float>(ipos) * __ldg(inv_freq + f), &sn, &cs);
        }
        const int dp = d < HALF ? d + HALF : d - HALF;
        const float sgn = d < HALF ? -1.f : 1.f;
        r[t] = row[d] * cs + sgn * row[dp] * sn;
        ss += r[t] * r[t];
      }
    }
    ss = warp_sum(ss);
    const float rs = 1.0f / sqrtf(ss / static_cast<float>(HD) + eps);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < DPL; ++t) {
      const int d = lane + 32 * t;
      if (d < HD) {
        float v = __ldg(w1 + d) * (r[t] * rs);
        if (w2 != nullptr) v *= __ldg(w2 + d);
        row[d] = v;
      }
    }
  };

  for (int64_t w = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; w < batch * num_heads;
       w += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t b = w / num_heads;
    const int h = static_cast<int>(w - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;
    // ---- the new tokens' queries
    for (int i = 0; i < MQ; ++i) {
      const int64_t base = (b * MQ + i) * ld + h * HD;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) sQ[i * LDS + d] = ld_qkv<QKV_BF16>(q_region, base + d);
      }
    }
    __syncwarp();
    for (int i = 0; i < MQ; ++i) {
      condition_row(sQ + i * LDS, q_pos0 + i - nm, q_ln_w, q_scale);
      __syncwarp();
    }
    float m_run[MQ], l_run[MQ], o[MQ][DPL];
#pragma unroll
    for (int i = 0; i < MQ; ++i) {
      m_run[i] = -INFINITY, l_run[i] = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) o[i][t] = 0.f;
    }
    // ---- keys / values tile by tile over all regions
    int region = 0, region_start = 0;  // sequence index of the region's first token
    for (int j0 = 0; j0 < total_tokens; j0 += KT) {
      const int tile = min(KT, total_tokens - j0);
      // stage raw k, v rows of the tile (a tile may straddle regions)
      int rr = region, rs = region_start;
      for (int jj = 0; jj < tile; ++jj) {
        const int j = j0 + jj;
        while (j >= rs + regions.tokens[rr]) {
          rs += regions.tokens[rr];
          ++rr;
        }
        const int64_t base = (b * regions.tokens[rr] + (j - rs)) * ld + h * HD;
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
          const int d = lane + 32 * t;
          if (d < HD) {
            sK[jj * LDS + d] = ld_qkv<QKV_BF16>(regions.ptr[rr], base + width + d);
            sV[jj * LDS + d] = ld_qkv<QKV_BF16>(regions.ptr[rr], base + 2 * width + d);
          }
        }
      }
      region = rr, region_start = rs;
      __syncwarp();
      for (int jj = 0; jj < tile; ++jj) {
        condition_row(sK + jj * LDS, j0 + jj - nm, k_ln_w, nullptr);
        __syncwarp();
      }
      // scores: lane = key of the tile
      const int j = j0 + lane;
      const bool key_ok = lane < tile && (j >= n_ctx || patch_mask == nullptr || patch_mask[b * n_ctx + j] == 0);
#pragma unroll
      for (int i = 0; i < MQ; ++i) {
        float s = -INFINITY;
        if (key_ok && j <= q_pos0 + i) {
          float acc = 0.f;
#pragma unroll 8
          for (int d = 0; d < HD; ++d) acc = fmaf(sQ[i * LDS + d], sK[lane * LDS + d], acc);
          s = acc;
        }
        const float m_new = fmaxf(m_run[i], warp_max(s));
        // a new token always sees itself, so m_new is finite from the tile that holds it on; before that every
        // probability of the tile is zero and the running state stays empty
        const float p = (s == -INFINITY || m_new == -INFINITY) ? 0.f : expf(s - m_new);
        const float corr = (m_run[i] == -INFINITY) ? 0.f : expf(m_run[i] - m_new);
        l_run[i] = l_run[i] * corr + warp_sum(p);
        m_run[i] = m_new;
        sP[i * KT + lane] = p;
#pragma unroll
        for (int t = 0; t < DPL; ++t) o[i][t] *= corr;
      }
      __syncwarp();
      for (int jj = 0; jj < tile; ++jj) {
#pragma unroll
        for (int i = 0; i < MQ; ++i) {
          const float p = sP[i * KT + jj];
#pragma unroll
          for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            if (d < HD) o[i][t] = fmaf(p, sV[jj * LDS + d], o[i][t]);
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < MQ; ++i) {
      const float inv = 1.0f / l_run[i];
      const int64_t row = b * MQ + i;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          const float v = o[i][t] * inv;
          const int c = h * HD + d;
          if constexpr (OUT == TSFMX_DT_F32) {
            reinterpret_cast<float*>(out)[row * width + c] = v;
          } else if constexpr (OUT == TSFMX_DT_BF16) {
            reinterpret_cast<__nv_bfloat16*>(out)[row * width + c] = __float2bfloat16_rn(v);
          } else {
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            reinterpret_cast<__nv_bfloat16*>(out)[row * 2 * width + c] = hi;
            reinterpret_cast<__nv_bfloat16*>(out)[row * 2 * width + width + c] = lo;
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ forecast finalize
// One block per series.  pf [(1 + flip) * B, ht, Q] (rows B.. = forecasts of the negated inputs), spread
// [(1 + flip) * B, hs, Q] or NULL, inputs [B, C] (positivity test: min over the context >= 0) -> out [B, horizon, Q].
__global__ void timesfm_forecast_finalize_kernel(const float* __restrict__ pf, const float* __restrict__ spread,
                                                 const float* __restrict__ inputs, int64_t batch, int context, int ht,
                                                 int hs, int nq, int horizon, int decode_index, int flip, int use_cq,
                                                 int positive, float* __restrict__ out) {
  const int64_t b = blockIdx.x;
  __shared__ float s_min[32];
  __shared__ int s_pos;
  if (positive) {
    float mn = INFINITY;
    for (int i = threadIdx.x; i < context; i += blockDim.x) mn = fminf(mn, inputs[b * context + i]);
    mn = -warp_max(-mn);
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = INFINITY;
      for (int i = 0; i < static_cast<int>((blockDim.x + 31) >> 5); ++i) m = fminf(m, s_min[i]);
      s_pos = m >= 0.f ? 1 : 0;
    }
    __syncthreads();
  }
  const bool clamp = positive && s_pos;
  const float* pf_p = pf + b * ht * nq;
  const float* pf_n = pf + (batch + b) * ht * nq;
  const float* sp_p = spread != nullptr ? spread + b * hs * nq : nullptr;
  const float* sp_n = spread != nullptr ? spread + (batch + b) * hs * nq : nullptr;
  // channel c of the combined tensor: (x[c] - x_neg[flip(c)]) / 2, flip(0) = 0, flip(c) = nq - c for c >= 1
  auto combined = [&](const float* pos, const float* neg, int t, int c) {
    const float v = pos[t * nq + c];
    if (!flip) return v;
    const int fc = c == 0 ? 0 : nq - c;
    return (v - neg[t * nq + fc]) / 2.f;
  };
  const int hq = use_cq ? min(horizon, hs) : 0;
  for (int i = threadIdx.x; i < horizon * nq; i += blockDim.x) {
    const int t = i / nq, c = i - t * nq;
    float v = combined(pf_p, pf_n, t, c);
    if (use_cq && t < hq && c >= 1 && c != decode_index) {
      v = combined(sp_p, sp_n, t, c) - combined(sp_p, sp_n, t, decode_index) + combined(pf_p, pf_n, t, decode_index);
    }
    if (clamp) v = fmaxf(v, 0.f);
    out[(b * horizon + t) * nq + c] = v;
  }
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_timesfm_patchify_continue(const float* x, int64_t x_series_stride, int64_t x_elem_stride,
                                               int64_t batch, int32_t patches, int32_t patch_len, float* state_n,
                                               float* state_mu, float* state_sigma, int32_t tokens_dtype, void* tokens,
                                               float* mu, float* sigma, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(x != nullptr && state_n != nullptr && state_mu != nullptr && state_sigma != nullptr &&
                    tokens != nullptr && mu != nullptr && sigma != nullptr,
                "timesfm_patchify_continue: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && patches > 0, "timesfm_patchify_continue: bad sizes");
  TSFMX_REQUIRE(tokens_dtype >= TSFMX_DT_F32 && tokens_dtype <= TSFMX_DT_BF16_SPLIT,
                "timesfm_patchify_continue: bad tokens_dtype");
  if (patch_len != 32) {
    set_error("timesfm_patchify_continue: patch_len %d unsupported (TimesFM 2.5 uses 32)", patch_len);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  const int threads = 256;
  const int64_t blocks = (batch * 32 + threads - 1) / threads;
  auto launch = [&](auto kern) {
    kern<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x, x_series_stride, x_elem_stride, batch, patches,
                                                                 state_n, state_mu, state_sigma, tokens, mu, sigma);
    return check_last_launch("timesfm_patchify_continue");
  };
  if (tokens_dtype == TSFMX_DT_F32) return launch(timesfm_patchify_continue_kernel<TSFMX_DT_F32>);
  if (tokens_dtype == TSFMX_DT_BF16) return launch(timesfm_patchify_continue_kernel<TSFMX_DT_BF16>);
  return launch(timesfm_patchify_continue_kernel<TSFMX_DT_BF16_SPLIT>);
}

extern "C" int tsfmx_timesfm_attention_decode(const void* const* region_ptrs, const int32_t* region_tokens,
                                              int32_t num_regions, int32_t qkv_dtype, int64_t batch, int32_t num_heads,
                                              int32_t head_dim, int32_t n_ctx, const uint8_t* patch_mask,
                                              const int32_t* num_masked, const float* inv_freq, const float* q_ln_w,
                                              const float* k_ln_w, const float* q_scale, float eps, int32_t out_dtype,
                                              void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(region_ptrs != nullptr && region_tokens != nullptr && out != nullptr && inv_freq != nullptr &&
                    q_ln_w != nullptr && k_ln_w != nullptr && q_scale != nullptr,
                "timesfm_attention_decode: NULL pointer");
  TSFMX_REQUIRE(num_regions >= 1 && num_regions <= MAX_REGIONS, "timesfm_attention_decode: 1..%d regions, got %d",
                MAX_REGIONS, num_regions);
  TSFMX_REQUIRE(batch >= 0 && num_heads > 0, "timesfm_attention_decode: bad sizes");
  TSFMX_REQUIRE(qkv_dtype == TSFMX_DT_F32 || qkv_dtype == TSFMX_DT_BF16,
                "timesfm_attention_decode: qkv must be f32 or bf16");
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "timesfm_attention_decode: bad out_dtype");
  KvRegions regions;
  regions.count = num_regions;
  for (int r = 0; r < num_regions; ++r) {
    TSFMX_REQUIRE(region_ptrs[r] != nullptr && region_tokens[r] > 0, "timesfm_attention_decode: empty region %d", r);
    regions.ptr[r] = region_ptrs[r];
    regions.tokens[r] = region_tokens[r];
  }
  TSFMX_REQUIRE(n_ctx >= 0 && (num_regions == 1 || n_ctx <= region_tokens[0]) && (patch_mask == nullptr || n_ctx > 0),
                "timesfm_attention_decode: n_ctx (%d) must be the padded-mask length of region 0", n_ctx);
  if (head_dim != 80 || region_tokens[num_regions - 1] != 4) {
    set_error("timesfm_attention_decode: head_dim %d / %d new tokens unsupported (TimesFM 2.5: 80, 128 / 32 = 4)",
              head_dim, region_tokens[num_regions - 1]);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  constexpr int HD = 80, MQ = 4;
  const int per_warp = ((MQ + 64) * (HD + 1) + MQ * 32) * 4;
  int total_tokens = 0;
  for (int r = 0; r < num_regions; ++r) total_tokens += region_tokens[r];
  int rope_rows = n_ctx + total_tokens;  // positions -n_ctx .. total_tokens - 1 (num_masked <= n_ctx)
  int rope_bytes = rope_rows * (HD / 2) * static_cast<int>(sizeof(float2));
  if (rope_bytes > 96 * 1024) rope_rows = 0, rope_bytes = 0;
  // 8 warps share one table where that still fits an SM's shared memory, else two blocks of 4 warps
  const int wpb = (rope_rows > 0 && 8 * per_warp + rope_bytes <= 224 * 1024) ? 8 : 4;
  const int smem = wpb * per_warp + rope_bytes;
  const int64_t total = batch * num_heads;
  const int64_t blocks = (total + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("timesfm_attention_decode: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
    kern<<<grid, wpb * 32, smem, stream>>>(regions, batch, num_heads, n_ctx, patch_mask, num_masked, inv_freq, q_ln_w,
                                            k_ln_w, q_scale, eps, rope_rows, out);
    return check_last_launch("timesfm_attention_decode");
  };
  if (qkv_dtype == TSFMX_DT_BF16) {
    if (out_dtype == TSFMX_DT_F32) return launch(timesfm_attention_decode_kernel<HD, MQ, 1, TSFMX_DT_F32>);
    if (out_dtype == TSFMX_DT_BF16) return launch(timesfm_attention_decode_kernel<HD, MQ, 1, TSFMX_DT_BF16>);
    return launch(timesfm_attention_decode_kernel<HD, MQ, 1, TSFMX_DT_BF16_SPLIT>);
  }
  if (out_dtype == TSFMX_DT_F32) return launch(timesfm_attention_decode_kernel<HD, MQ, 0, TSFMX_DT_F32>);
  if (out_dtype == TSFMX_DT_BF16) return launch(timesfm_attention_decode_kernel<HD, MQ, 0, TSFMX_DT_BF16>);
  return launch(timesfm_attention_decode_kernel<HD, MQ, 0, TSFMX_DT_BF16_SPLIT>);
}

extern "C" int tsfmx_timesfm_forecast_finalize(const float* pf, const float* spread, const float* inputs, int64_t batch,
                                               int32_t context, int32_t pf_steps, int32_t spread_steps,
                                               int32_t num_outputs, int32_t horizon, int32_t decode_index, int32_t flip,
                                               int32_t use_continuous_quantile_head, int32_t infer_is_positive,
                                               float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(pf != nullptr && out != nullptr, "timesfm_forecast_finalize: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && horizon > 0 && horizon <= pf_steps && num_outputs > 1,
                "timesfm_forecast_finalize: horizon %d must be in [1, %d]", horizon, pf_steps);
  TSFMX_REQUIRE(decode_index >= 0 && decode_index < num_outputs, "timesfm_forecast_finalize: bad decode_index");
  TSFMX_REQUIRE(!use_continuous_quantile_head || (spread != nullptr && spread_steps > 0),
                "timesfm_forecast_finalize: the continuous quantile head needs the quantile spread");
  TSFMX_REQUIRE(!infer_is_positive || (inputs != nullptr && context > 0),
                "timesfm_forecast_finalize: infer_is_positive needs the inputs");
  if (batch == 0) return TSFMX_OK;
  timesfm_forecast_finalize_kernel<<<static_cast<unsigned>(batch), 256, 0, stream>>>(
      pf, spread, inputs, batch, context, pf_steps, spread_steps, num_outputs, horizon, decode_index, flip ? 1 : 0,
      use_continuous_quantile_head ? 1 : 0, infer_is_positive ? 1 : 0, out);
  return check_last_launch("timesfm_forecast_finalize");
}

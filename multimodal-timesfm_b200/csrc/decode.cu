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
//                                      fly in 32-key tiles (one lane per key) with an online softmax, so any
//                                      context length fits.
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

// One warp per (series, head).  MQ = 4 new tokens (queries) = tokens of the last region.
//
// Work layout (second version; the first one conditioned every key row with the lanes spread over head_dim - a warp
// reduction and ~60 instructions per key row - and staged V in shared memory although every V element is used once):
//   * queries: conditioned with lanes over head_dim (4 rows only), pre-multiplied by k_ln_w, stored TRANSPOSED
//     (sQT[d][4]) so that one 16-byte shared-memory load feeds the four dot products of a key;
//   * keys: a tile of 32 raw key rows is staged through shared memory (coalesced global reads), then LANE j owns key
//     j: it pulls its row into registers, rotates it with (cos, sin) from a global table (L1-resident, built once per
//     inv_freq by tsfmx_rope_table), accumulates its own sum of squares (no shuffles) and keeps 1 / rms as a scalar
//     that multiplies the score instead of the 80 elements;
//   * values: read straight from global memory in the P.V loop (lanes over head_dim, 160 contiguous bytes per row).
template <int HD, int MQ, int QKV_BF16, int OUT>
__global__ void __launch_bounds__(256) timesfm_attention_decode_kernel(
    KvRegions regions, int64_t batch, int num_heads, int n_ctx, const uint8_t* __restrict__ patch_mask,
    const int32_t* __restrict__ num_masked, const float2* __restrict__ rope, int rope_len,
    const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w, const float* __restrict__ k_ln_w,
    const float* __restrict__ q_scale, float eps, void* __restrict__ out) {
  static_assert(MQ == 4, "the transposed query tile is read as float4");
  constexpr int DPL = (HD + 31) / 32;
  constexpr int HALF = HD / 2;
  constexpr int LDS = HD + 1;
  constexpr int KT = 32;  // keys per tile
  extern __shared__ __align__(16) float smem_dec[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PER_WARP = HD * MQ + MQ * KT + MQ * LDS;  // sQT | sP | sQ (row-major scratch); a multiple of 4 floats
  static_assert(PER_WARP % 4 == 0, "float4 alignment of the per-warp tiles");
  float* sQT = smem_dec + warp * PER_WARP;
  float* sP = sQT + HD * MQ;   // [KT][MQ]
  float* sQ = sP + MQ * KT;    // [MQ][LDS]
  const int width = num_heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  int total_tokens = 0;
  for (int r = 0; r < regions.count; ++r) total_tokens += regions.tokens[r];
  const int q_pos0 = total_tokens - MQ;  // sequence index of the first new token
  const void* q_region = regions.ptr[regions.count - 1];

  auto rotation = [&](int ipos, int f, float& cs, float& sn) {  // the launcher checks that the table covers |ipos|
    const int ap = ipos < 0 ? -ipos : ipos;
    const float2 e = __ldg(rope + static_cast<int64_t>(ap) * HALF + f);
    cs = e.x, sn = ipos < 0 ? -e.y : e.y;
  };

  for (int64_t w = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; w < batch * num_heads;
       w += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t b = w / num_heads;
    const int h = static_cast<int>(w - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;
    // ---- the new tokens' queries: RoPE, RMSNorm, q_ln * q_scale, and k_ln folded in; lanes over head_dim
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
      float r[DPL];
      float ss = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        r[t] = 0.f;
        if (d < HD) {
          const int f = d < HALF ? d : d - HALF;
          float sn, cs;
          rotation(q_pos0 + i - nm, f, cs, sn);
          const int dp = d < HALF ? d + HALF : d - HALF;
          const float sgn = d < HALF ? -1.f : 1.f;
          r[t] = sQ[i * LDS + d] * cs + sgn * sQ[i * LDS + dp] * sn;
          ss += r[t] * r[t];
        }
      }
      ss = warp_sum(ss);
      const float rs = 1.0f / sqrtf(ss / static_cast<float>(HD) + eps);
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) sQT[d * MQ + i] = __ldg(q_ln_w + d) * (r[t] * rs) * __ldg(q_scale + d) * __ldg(k_ln_w + d);
      }
    }
    __syncwarp();
    float m_run[MQ], l_run[MQ], o[MQ][DPL];
#pragma unroll
    for (int i = 0; i < MQ; ++i) {
      m_run[i] = -INFINITY, l_run[i] = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) o[i][t] = 0.f;
    }
    // ---- keys / values tile by tile over all regions.  Region 0 holds n0 tokens per series, every later region MQ
    //      (one per decode step; checked by the launcher), so a sequence index maps to (region, row) without a search
    const int n0 = regions.tokens[0];
    auto row_base = [&](int j, const void*& ptr) -> int64_t {  // element index of q[0] of head h of sequence token j
      const int r = j < n0 ? 0 : 1 + (j - n0) / MQ;
      const int local = j < n0 ? j : (j - n0) % MQ;
      ptr = regions.ptr[r];
      return (b * (r == 0 ? n0 : MQ) + local) * ld + h * HD;
    };
    for (int j0 = 0; j0 < total_tokens; j0 += KT) {
      const int tile = min(KT, total_tokens - j0);
      // lane = key: the lane pulls its own raw key row (160 / 320 contiguous bytes) and its row of the rotation table
      // straight into registers - all loads independent, nothing staged in shared memory, so the SM keeps its L1 for
      // the table - rotates, accumulates its own sum of squares and the four dot products
      const int j = j0 + lane;
      const bool key_ok = lane < tile && (j >= n_ctx || patch_mask == nullptr || patch_mask[b * n_ctx + j] == 0);
      float s4[MQ] = {0.f, 0.f, 0.f, 0.f};
      float krs = 0.f;
      if (lane < tile) {
        float kf[HD];
        {
          const void* ptr;
          const int64_t base = row_base(j, ptr) + width;
          if constexpr (QKV_BF16) {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ptr) + base);
            uint4 raw[HD / 8];
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) raw[c] = src[c];
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              const uint32_t wv[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                kf[8 * c + 2 * k] = __uint_as_float(wv[k] << 16);
                kf[8 * c + 2 * k + 1] = __uint_as_float(wv[k] & 0xffff0000u);
              }
            }
          } else {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ptr) + base);
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
              const float4 v = src[c];
              kf[4 * c] = v.x, kf[4 * c + 1] = v.y, kf[4 * c + 2] = v.z, kf[4 * c + 3] = v.w;
            }
          }
        }
        const int ipos = j - nm;
        const int ap = ipos < 0 ? -ipos : ipos;
        const float sgn = ipos < 0 ? -1.f : 1.f;  // sin(-x) = -sin(x): the table holds non-negative positions
        const float4* trow = reinterpret_cast<const float4*>(rope + static_cast<int64_t>(ap) * HALF);
        float ss = 0.f;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          float4 t[HALF / 4];
#pragma unroll
          for (int g = 0; g < HALF / 4; ++g) t[g] = __ldg(trow + part * (HALF / 4) + g);
#pragma unroll
          for (int g = 0; g < HALF / 4; ++g) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int f = part * (HALF / 2) + 2 * g + e;
              const float cs = e == 0 ? t[g].x : t[g].z;
              const float sn = sgn * (e == 0 ? t[g].y : t[g].w);
              const float a = kf[f], c = kf[f + HALF];
              const float r1 = a * cs - c * sn, r2 = c * cs + a * sn;
              ss = fmaf(r1, r1, fmaf(r2, r2, ss));
              const float4 q1 = *reinterpret_cast<const float4*>(sQT + f * MQ);
              const float4 q2 = *reinterpret_cast<const float4*>(sQT + (f + HALF) * MQ);
              s4[0] = fmaf(q1.x, r1, fmaf(q2.x, r2, s4[0]));
              s4[1] = fmaf(q1.y, r1, fmaf(q2.y, r2, s4[1]));
              s4[2] = fmaf(q1.z, r1, fmaf(q2.z, r2, s4[2]));
              s4[3] = fmaf(q1.w, r1, fmaf(q2.w, r2, s4[3]));
            }
          }
        }
        krs = 1.0f / sqrtf(ss / static_cast<float>(HD) + eps);
      }
      float p4[MQ];
#pragma unroll
      for (int i = 0; i < MQ; ++i) {
        const float s = (key_ok && j <= q_pos0 + i) ? s4[i] * krs : -INFINITY;
        const float m_new = fmaxf(m_run[i], warp_max(s));
        // a new token always sees itself, so m_new is finite from the tile that holds it on; before that every
        // probability of the tile is zero and the running state stays empty
        const float p = (s == -INFINITY || m_new == -INFINITY) ? 0.f : expf(s - m_new);
        const float corr = (m_run[i] == -INFINITY) ? 0.f : expf(m_run[i] - m_new);
        l_run[i] = l_run[i] * corr + warp_sum(p);
        m_run[i] = m_new;
        p4[i] = p;
#pragma unroll
        for (int t = 0; t < DPL; ++t) o[i][t] *= corr;
      }
      *reinterpret_cast<float4*>(sP + lane * MQ) = make_float4(p4[0], p4[1], p4[2], p4[3]);
      __syncwarp();
      // P.V with the value rows read from global memory (each element is used exactly once); unrolled so that the loads
      // of several key rows are in flight together
#pragma unroll 8
      for (int jj = 0; jj < tile; ++jj) {
        const void* ptr;
        const int64_t base = row_base(j0 + jj, ptr) + 2 * width;
        const float4 p = *reinterpret_cast<const float4*>(sP + jj * MQ);
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
          const int d = lane + 32 * t;
          if (d < HD) {
            const float v = ld_qkv<QKV_BF16>(ptr, base + d);
            o[0][t] = fmaf(p.x, v, o[0][t]);
            o[1][t] = fmaf(p.y, v, o[1][t]);
            o[2][t] = fmaf(p.z, v, o[2][t]);
            o[3][t] = fmaf(p.w, v, o[3][t]);
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
                                              const int32_t* num_masked, const float* rope_table, int32_t rope_len,
                                              const float* inv_freq, const float* q_ln_w, const float* k_ln_w,
                                              const float* q_scale, float eps, int32_t out_dtype, void* out,
                                              void* stream_) {
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
  TSFMX_REQUIRE(rope_len > 0 && rope_table != nullptr, "timesfm_attention_decode: rotation table missing");
  KvRegions regions;
  regions.count = num_regions;
  for (int r = 0; r < num_regions; ++r) {
    TSFMX_REQUIRE(region_ptrs[r] != nullptr && region_tokens[r] > 0, "timesfm_attention_decode: empty region %d", r);
    regions.ptr[r] = region_ptrs[r];
    regions.tokens[r] = region_tokens[r];
  }
  TSFMX_REQUIRE(n_ctx >= 0 && (num_regions == 1 || n_ctx <= region_tokens[0]) && (patch_mask == nullptr || n_ctx > 0),
                "timesfm_attention_decode: n_ctx (%d) must be the padded-mask length of region 0", n_ctx);
  for (int r = 1; r < num_regions; ++r)
    TSFMX_REQUIRE(region_tokens[r] == 4, "timesfm_attention_decode: region %d holds %d tokens per series; every region after "
                  "the prefill's must hold the 4 tokens of one decode step", r, region_tokens[r]);
  for (int r = 0; r < num_regions; ++r)
    TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(region_ptrs[r]) % 16 == 0, "timesfm_attention_decode: region %d is not 16-byte aligned", r);
  if (head_dim != 80 || region_tokens[num_regions - 1] != 4) {
    set_error("timesfm_attention_decode: head_dim %d / %d new tokens unsupported (TimesFM 2.5: 80, 128 / 32 = 4)",
              head_dim, region_tokens[num_regions - 1]);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  constexpr int HD = 80, MQ = 4;
  const int per_warp = (HD * MQ + MQ * 32 + MQ * (HD + 1)) * 4;  // 3.1 KB: the SM keeps its L1 for the rotation table
  const int wpb = 8;
  const int smem = wpb * per_warp;
  int total_tokens = 0;
  for (int r = 0; r < num_regions; ++r) total_tokens += region_tokens[r];
  TSFMX_REQUIRE(rope_len >= total_tokens && rope_len >= n_ctx,
                "timesfm_attention_decode: the rotation table (%d positions) must cover %d tokens", rope_len,
                total_tokens > n_ctx ? total_tokens : n_ctx);
  const int64_t total = batch * num_heads;
  const int64_t blocks = (total + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  auto launch = [&](auto kern) -> int {
    kern<<<grid, wpb * 32, smem, stream>>>(regions, batch, num_heads, n_ctx, patch_mask, num_masked,
                                            reinterpret_cast<const float2*>(rope_table), rope_len, inv_freq, q_ln_w, k_ln_w,
                                            q_scale, eps, out);
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

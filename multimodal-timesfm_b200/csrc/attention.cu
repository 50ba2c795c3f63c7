// TimesFM 2.5 attention core (everything between qkv_proj and the output projection).
//
// Sequences on this path are tiny (N = context / 32 patches: 16 at ctx 512, 64 at ctx 2048) and the
// attention FLOPs are ~0.2 % of a layer, so one warp owns one (series, head) pair, keeps the
// conditioned Q/K/V of that pair in shared memory and does the whole flash-style chain
// RoPE -> RMSNorm(q), RMSNorm(k) -> per-dim query scale -> masked softmax -> P.V in fp32 without
// ever materialising scores in HBM.
//
// Follows upstream timesfm MultiHeadAttention (HF twin modeling_timesfm2_5.py:304-346):
//   position = n - num_masked[b]; rotate-half RoPE; RMSNorm over head_dim AFTER RoPE;
//   q *= softplus(per_dim_scale) * 1.442695041 / sqrt(hd)  (precomputed by the host as q_scale);
//   mask = causal & key-not-padded; softmax scale 1.0.
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

template <int QKV_BF16>
__device__ __forceinline__ float load_qkv(const void* qkv, int64_t idx) {
  if constexpr (QKV_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[idx]);
  else return reinterpret_cast<const float*>(qkv)[idx];
}

template <int OUT>
__device__ __forceinline__ void store_out(void* out, int64_t row, int width, int c, float v) {
  if constexpr (OUT == TSFMX_DT_F32) {
    reinterpret_cast<float*>(out)[row * width + c] = v;
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    reinterpret_cast<__nv_bfloat16*>(out)[row * width + c] = __float2bfloat16_rn(v);
  } else {
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    reinterpret_cast<__nv_bfloat16*>(out)[row * 2 * width + c] = h;
    reinterpret_cast<__nv_bfloat16*>(out)[row * 2 * width + width + c] = l;
  }
}

// HD = head_dim (80), DPL = ceil(HD / 32) elements per lane
template <int HD, int QKV_BF16, int OUT>
__global__ void timesfm_attention_kernel(const void* __restrict__ qkv, int64_t batch, int num_patches, int num_heads,
                                         const uint8_t* __restrict__ patch_mask, const int32_t* __restrict__ num_masked,
                                         const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w,
                                         const float* __restrict__ k_ln_w, const float* __restrict__ q_scale, float eps,
                                         void* out) {
  constexpr int DPL = (HD + 31) / 32;
  constexpr int HALF = HD / 2;
  constexpr int LDS = HD + 1;  // padded row: key-parallel dot products are bank-conflict free
  extern __shared__ float smem[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = num_patches;
  const int per_warp = 3 * N * LDS + N;
  float* sQ = smem + warp * per_warp;
  float* sK = sQ + N * LDS;
  float* sV = sK + N * LDS;
  float* sP = sV + N * LDS;
  const int width = num_heads * HD;        // 1280
  const int64_t qkv_ld = 3 * static_cast<int64_t>(width);
  const int64_t total = batch * num_heads;

  for (int64_t w = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; w < total;
       w += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t b = w / num_heads;
    const int h = static_cast<int>(w - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;
    const uint8_t* pm = patch_mask != nullptr ? patch_mask + b * N : nullptr;

    // ---- stage raw q, k, v of this (series, head)
    for (int n = 0; n < N; ++n) {
      const int64_t base = (b * N + n) * qkv_ld + h * HD;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          sQ[n * LDS + d] = load_qkv<QKV_BF16>(qkv, base + d);
          sK[n * LDS + d] = load_qkv<QKV_BF16>(qkv, base + width + d);
          sV[n * LDS + d] = load_qkv<QKV_BF16>(qkv, base + 2 * width + d);
        }
      }
    }
    __syncwarp();
    // ---- RoPE + RMSNorm (+ per-dim scale on q), row by row, in place
    for (int n = 0; n < N; ++n) {
      const float pos = static_cast<float>(n - nm);
      float qr[DPL], kr[DPL];
      float qss = 0.f, kss = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        qr[t] = 0.f, kr[t] = 0.f;
        if (d < HD) {
          const int f = d < HALF ? d : d - HALF;
          const float ang = pos * __ldg(inv_freq + f);
          float sn, cs;
          sincosf(ang, &sn, &cs);
          const int dp = d < HALF ? d + HALF : d - HALF;
          const float sgn = d < HALF ? -1.f : 1.f;
          qr[t] = sQ[n * LDS + d] * cs + sgn * sQ[n * LDS + dp] * sn;
          kr[t] = sK[n * LDS + d] * cs + sgn * sK[n * LDS + dp] * sn;
          qss += qr[t] * qr[t];
          kss += kr[t] * kr[t];
        }
      }
      qss = warp_sum(qss);
      kss = warp_sum(kss);
      const float qrs = 1.0f / sqrtf(qss / static_cast<float>(HD) + eps);
      const float krs = 1.0f / sqrtf(kss / static_cast<float>(HD) + eps);
      __syncwarp();  // all lanes have read the un-rotated row
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          sQ[n * LDS + d] = __ldg(q_ln_w + d) * (qr[t] * qrs) * __ldg(q_scale + d);
          sK[n * LDS + d] = __ldg(k_ln_w + d) * (kr[t] * krs);
        }
      }
    }
    __syncwarp();
    // ---- per query row: scores -> masked softmax -> P.V
    for (int i = 0; i < N; ++i) {
      float mx = -INFINITY;
      bool any = false;
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j < N) {
          const bool allowed = (j <= i) && (pm == nullptr || pm[j] == 0);
          if (allowed) {
            float acc = 0.f;
#pragma unroll 8
            for (int d = 0; d < HD; ++d) acc = fmaf(sQ[i * LDS + d], sK[j * LDS + d], acc);
            s = acc;
            any = true;
          }
          sP[j] = s;
        }
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      const bool row_has_key = __any_sync(0xffffffffu, any);
      float sum = 0.f;
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        if (j < N) {
          // a row with every key masked gets uniform weights over all N keys (additive finfo.min mask)
          const float p = row_has_key ? (sP[j] == -INFINITY ? 0.f : expf(sP[j] - mx)) : 1.f;
          sP[j] = p;
          sum += p;
        }
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      __syncwarp();
      float o[DPL];
#pragma unroll
      for (int t = 0; t < DPL; ++t) o[t] = 0.f;
      const int jend = row_has_key ? i + 1 : N;
      for (int j = 0; j < jend; ++j) {
        const float p = sP[j];
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
          const int d = lane + 32 * t;
          if (d < HD) o[t] = fmaf(p, sV[j * LDS + d], o[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) store_out<OUT>(out, b * N + i, width, h * HD + d, o[t] * inv);
      }
      __syncwarp();
    }
  }
}

template <int HD, int QKV_BF16>
int launch_attention(const void* qkv, int64_t batch, int N, int H, const uint8_t* pm, const int32_t* nm,
                     const float* inv_freq, const float* qw, const float* kw, const float* qs, float eps, int out_dtype,
                     void* out, cudaStream_t stream) {
  const int per_warp_bytes = (3 * N * (HD + 1) + N) * 4;
  int wpb = (96 * 1024) / per_warp_bytes;
  if (wpb > 4) wpb = 4;
  if (wpb < 1) {
    set_error("timesfm_attention: %d patches need %d bytes of shared memory per warp; unsupported", N, per_warp_bytes);
    return TSFMX_ERR_UNSUPPORTED;
  }
  const int smem = wpb * per_warp_bytes;
  const int64_t total = batch * H;
  const int64_t blocks = (total + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("timesfm_attention: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
    kern<<<grid, wpb * 32, smem, stream>>>(qkv, batch, N, H, pm, nm, inv_freq, qw, kw, qs, eps, out);
    return check_last_launch("timesfm_attention");
  };
  if (out_dtype == TSFMX_DT_F32) return launch(timesfm_attention_kernel<HD, QKV_BF16, TSFMX_DT_F32>);
  if (out_dtype == TSFMX_DT_BF16) return launch(timesfm_attention_kernel<HD, QKV_BF16, TSFMX_DT_BF16>);
  return launch(timesfm_attention_kernel<HD, QKV_BF16, TSFMX_DT_BF16_SPLIT>);
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_timesfm_attention(const void* qkv, int32_t qkv_dtype, int64_t batch, int32_t num_patches,
                                       int32_t num_heads, int32_t head_dim, const uint8_t* patch_mask,
                                       const int32_t* num_masked, const float* inv_freq, const float* q_ln_w,
                                       const float* k_ln_w, const float* q_scale, float eps, int32_t out_dtype,
                                       void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && out != nullptr && inv_freq != nullptr && q_ln_w != nullptr && k_ln_w != nullptr &&
                    q_scale != nullptr,
                "timesfm_attention: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && num_patches > 0 && num_heads > 0, "timesfm_attention: bad sizes");
  TSFMX_REQUIRE(qkv_dtype == TSFMX_DT_F32 || qkv_dtype == TSFMX_DT_BF16, "timesfm_attention: qkv must be f32 or bf16");
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "timesfm_attention: bad out_dtype");
  if (head_dim != 80) {
    set_error("timesfm_attention: head_dim %d unsupported (TimesFM 2.5 uses 80)", head_dim);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  if (qkv_dtype == TSFMX_DT_BF16)
    return launch_attention<80, 1>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w, k_ln_w,
                                   q_scale, eps, out_dtype, out, stream);
  return launch_attention<80, 0>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w, k_ln_w,
                                 q_scale, eps, out_dtype, out, stream);
}

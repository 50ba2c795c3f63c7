// Chronos-2 specific stages (reference tsfmx/tsfm/chronos.py:62-169 -> upstream Chronos2Model.encoder,
// instance_norm.inverse):
//
//   encoder_attention   time self-attention core of an encoder block: RoPE(q, k) with positions arange(T),
//                       NO 1/sqrt(d) scaling, bidirectional, additive key-padding mask, fp32 softmax.
//                       One CTA per (series, head); the head's q/k/v (T <= 193 tokens x 64) live in shared
//                       memory, warps split the query rows.
//   chronos2_finalize   output head epilogue: (B*np, 21*16) patch-major quantile predictions ->
//                       sinh(x) * scale + loc, reordered to (B, horizon, 21)  (chronos.py:159-169)
//
// The group self-attention of a block needs no kernel: with group_ids = arange(B) (chronos.py:117) it is
// exactly h + (W_o W_v) RMSNorm(h), i.e. one tcgen05 GEMM on a pre-multiplied 768x768 weight.
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

__device__ __forceinline__ float ld_any(const void* base, int dtype, int64_t idx) {
  return dtype == TSFMX_DT_F32 ? reinterpret_cast<const float*>(base)[idx]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

template <int HD, int OUT>
__global__ void __launch_bounds__(128) encoder_attention_kernel(const void* __restrict__ qkv, int qkv_dtype,
                                                                 int seq, int num_heads,
                                                                 const uint8_t* __restrict__ key_mask,
                                                                 const float* __restrict__ inv_freq, void* out) {
  constexpr int HALF = HD / 2;
  constexpr int LD = HD + 1;
  constexpr int DPL = HD / 32;
  extern __shared__ float smem[];
  const int T = seq;
  float* sQ = smem;
  float* sK = sQ + T * LD;
  float* sV = sK + T * LD;
  float* sP = sV + T * LD;  // [4 warps][T]
  const int b = blockIdx.x / num_heads, h = blockIdx.x - b * num_heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = num_heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const int64_t base = static_cast<int64_t>(b) * T * ld + h * HD;

  // stage v, and q / k with the rotary embedding applied (rotate-half, position = token index)
  for (int i = threadIdx.x; i < T * HALF; i += blockDim.x) {
    const int t = i / HALF, d = i - t * HALF;
    float sn, cs;
    sincosf(static_cast<float>(t) * __ldg(inv_freq + d), &sn, &cs);
    const int64_t row = base + t * ld;
    const float q1 = ld_any(qkv, qkv_dtype, row + d), q2 = ld_any(qkv, qkv_dtype, row + d + HALF);
    const float k1 = ld_any(qkv, qkv_dtype, row + width + d), k2 = ld_any(qkv, qkv_dtype, row + width + d + HALF);
    sQ[t * LD + d] = q1 * cs - q2 * sn;
    sQ[t * LD + d + HALF] = q2 * cs + q1 * sn;
    sK[t * LD + d] = k1 * cs - k2 * sn;
    sK[t * LD + d + HALF] = k2 * cs + k1 * sn;
    sV[t * LD + d] = ld_any(qkv, qkv_dtype, row + 2 * width + d);
    sV[t * LD + d + HALF] = ld_any(qkv, qkv_dtype, row + 2 * width + d + HALF);
  }
  __syncthreads();
  const uint8_t* km = key_mask != nullptr ? key_mask + static_cast<int64_t>(b) * T : nullptr;
  float* p = sP + warp * T;
  for (int i = warp; i < T; i += 4) {
    float mx = -INFINITY;
    bool any = false;
    for (int j0 = 0; j0 < T; j0 += 32) {
      const int j = j0 + lane;
      float s = -INFINITY;
      if (j < T && (km == nullptr || km[j] != 0)) {
        float acc = 0.f;
#pragma unroll 8
        for (int d = 0; d < HD; ++d) acc = fmaf(sQ[i * LD + d], sK[j * LD + d], acc);
        s = acc;
        any = true;
      }
      if (j < T) p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    const bool has_key = __any_sync(0xffffffffu, any);
    float sum = 0.f;
    for (int j0 = 0; j0 < T; j0 += 32) {
      const int j = j0 + lane;
      if (j < T) {
        // every key masked: the additive finfo.min mask of the reference yields uniform weights
        const float e = has_key ? (p[j] == -INFINITY ? 0.f : expf(p[j] - mx)) : 1.f;
        p[j] = e;
        sum += e;
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float o[DPL];
#pragma unroll
    for (int t = 0; t < DPL; ++t) o[t] = 0.f;
    for (int j = 0; j < T; ++j) {
      const float pj = p[j];
#pragma unroll
      for (int t = 0; t < DPL; ++t) o[t] = fmaf(pj, sV[j * LD + lane + 32 * t], o[t]);
    }
    const int64_t orow = static_cast<int64_t>(b) * T + i;
#pragma unroll
    for (int t = 0; t < DPL; ++t) {
      const int c = h * HD + lane + 32 * t;
      const float v = o[t] * inv;
      if constexpr (OUT == TSFMX_DT_F32) {
        reinterpret_cast<float*>(out)[orow * width + c] = v;
      } else if constexpr (OUT == TSFMX_DT_BF16) {
        reinterpret_cast<__nv_bfloat16*>(out)[orow * width + c] = __float2bfloat16_rn(v);
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out) + orow * 2 * width;
        split_bf16(v, op[c], op[width + c]);
      }
    }
    __syncwarp();
  }
}

__global__ void chronos2_finalize_kernel(const float* __restrict__ preds, int64_t batch, int num_patches_used,
                                         int num_quantiles, int patch, int horizon, int use_arcsinh,
                                         const float* __restrict__ loc, const float* __restrict__ scale,
                                         float* __restrict__ out) {
  const int64_t total = batch * horizon * num_quantiles;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % num_quantiles);
    const int64_t bt = i / num_quantiles;
    const int t = static_cast<int>(bt % horizon);
    const int64_t b = bt / horizon;
    const int pi = t / patch, po = t - pi * patch;
    const float x = preds[(b * num_patches_used + pi) * (num_quantiles * patch) + q * patch + po];
    const float y = use_arcsinh ? sinhf(x) : x;
    out[i] = __fadd_rn(__fmul_rn(y, __ldg(scale + b)), __ldg(loc + b));
  }
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_encoder_attention(const void* qkv, int32_t qkv_dtype, int64_t batch, int32_t seq,
                                       int32_t num_heads, int32_t head_dim, const uint8_t* key_mask,
                                       const float* inv_freq, int32_t out_dtype, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && out != nullptr && inv_freq != nullptr, "encoder_attention: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && seq > 0 && num_heads > 0, "encoder_attention: bad sizes");
  TSFMX_REQUIRE(qkv_dtype == TSFMX_DT_F32 || qkv_dtype == TSFMX_DT_BF16, "encoder_attention: qkv must be f32 or bf16");
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "encoder_attention: bad out_dtype");
  TSFMX_REQUIRE(batch * num_heads < (int64_t(1) << 31), "encoder_attention: too many (series, head) pairs");
  if (head_dim != 64) {
    set_error("encoder_attention: head_dim %d unsupported (Chronos-2 uses 64)", head_dim);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  const int smem = (3 * seq * 65 + 4 * seq) * 4;
  if (smem > 220 * 1024) {
    set_error("encoder_attention: %d tokens need %d bytes of shared memory; unsupported", seq, smem);
    return TSFMX_ERR_UNSUPPORTED;
  }
  const int grid = static_cast<int>(batch * num_heads);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("encoder_attention: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
    kern<<<grid, 128, smem, stream>>>(qkv, qkv_dtype, seq, num_heads, key_mask, inv_freq, out);
    return check_last_launch("encoder_attention");
  };
  if (out_dtype == TSFMX_DT_F32) return launch(encoder_attention_kernel<64, TSFMX_DT_F32>);
  if (out_dtype == TSFMX_DT_BF16) return launch(encoder_attention_kernel<64, TSFMX_DT_BF16>);
  return launch(encoder_attention_kernel<64, TSFMX_DT_BF16_SPLIT>);
}

extern "C" int tsfmx_chronos2_finalize(const float* preds, int64_t batch, int32_t num_patches_used,
                                       int32_t num_quantiles, int32_t patch, int32_t horizon, int32_t use_arcsinh,
                                       const float* loc, const float* scale, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(preds != nullptr && loc != nullptr && scale != nullptr && out != nullptr,
                "chronos2_finalize: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && num_patches_used > 0 && num_quantiles > 0 && patch > 0 && horizon > 0,
                "chronos2_finalize: bad sizes");
  TSFMX_REQUIRE(horizon <= num_patches_used * patch, "chronos2_finalize: horizon (%d) exceeds %d patches of %d", horizon,
                num_patches_used, patch);
  const int64_t total = batch * horizon * num_quantiles;
  if (total == 0) return TSFMX_OK;
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
  chronos2_finalize_kernel<<<grid, 256, 0, stream>>>(preds, batch, num_patches_used, num_quantiles, patch, horizon,
                                                     use_arcsinh, loc, scale, out);
  return check_last_launch("chronos2_finalize");
}

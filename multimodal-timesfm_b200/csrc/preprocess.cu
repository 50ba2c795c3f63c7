// HBM-bound stages of the forecast path: every series is read once, its
// statistics are reduced in registers / warp shuffles, and the model-ready
// tokens are written once.  One warp owns one series; all global traffic is
// 16-byte vectorised and streaming (L1::no_allocate).
//
//   timesfm_patchify_norm   reference tsfmx/tsfm/timesfm.py:53-73
//   chronos2_patchify_norm  reference tsfmx/tsfm/chronos.py:48-52 (-> Chronos2Model._prepare_patched_context)
//   chronos_t5_tokenize     north-star item (upstream MeanScaleUniformBins)
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

constexpr int WARPS = 8;  // warps (= series in flight) per block

// ----------------------------------------------------------------------------------------
// TimesFM 2.5: patches of 32; lane l of a warp owns float4 #(l + 32 j) of a 512-element chunk,
// i.e. 8 consecutive lanes own one patch and a warp covers 16 patches per chunk.
// ----------------------------------------------------------------------------------------
template <int OUT>
__device__ __forceinline__ void store_token_quad(void* tokens, int64_t token_row, int q, const float (&val)[4],
                                                 const float (&msk)[4]) {
  // token row layout: [32 normalised values | 32 mask flags]; q = quad index inside the patch (0..7)
  if constexpr (OUT == TSFMX_DT_F32) {
    float* row = reinterpret_cast<float*>(tokens) + token_row * 64;
    st_stream_f4(row + 4 * q, make_float4(val[0], val[1], val[2], val[3]));
    st_stream_f4(row + 32 + 4 * q, make_float4(msk[0], msk[1], msk[2], msk[3]));
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(tokens) + token_row * 64;
    st_stream_u2(row + 4 * q, make_uint2(pack_bf16x2(val[0], val[1]), pack_bf16x2(val[2], val[3])));
    st_stream_u2(row + 32 + 4 * q, make_uint2(pack_bf16x2(msk[0], msk[1]), pack_bf16x2(msk[2], msk[3])));
  } else {
    // split: [hi(64) | lo(64)]
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(tokens) + token_row * 128;
    uint2 h, l;
    split_bf16x2(val[0], val[1], h.x, l.x);
    split_bf16x2(val[2], val[3], h.y, l.y);
    st_stream_u2(row + 4 * q, h);
    st_stream_u2(row + 64 + 4 * q, l);
    st_stream_u2(row + 32 + 4 * q, make_uint2(pack_bf16x2(msk[0], msk[1]), pack_bf16x2(msk[2], msk[3])));
    st_stream_u2(row + 96 + 4 * q, make_uint2(0u, 0u));
  }
}

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <int OUT>
__global__ void __launch_bounds__(WARPS * 32) timesfm_patchify_norm_kernel(
    const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t batch, int context, void* tokens,
    float* __restrict__ mu_out, float* __restrict__ sigma_out, uint8_t* __restrict__ patch_mask_out,
    int32_t* __restrict__ num_masked_out) {
  __shared__ float s_inc[WARPS][16][4];  // per patch of the chunk: inc_n, inc_mu, inc_sigma
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, q = lane & 7;
  const int num_patches = context >> 5;

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    const uint8_t* mr = mask + b * context;
    float run_n = 0.f, run_mu = 0.f, run_sigma = 0.f;
    int masked_patches = 0;

    for (int c0 = 0; c0 < context; c0 += 512) {
      float4 v[4];
      uint32_t mk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = c0 + 4 * (lane + 32 * j);
        if (e < context) {
          v[j] = ld_stream_f4(xr + e);
          mk[j] = ld_stream_u32(mr + e);
        } else {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          mk[j] = 0x01010101u;
        }
      }
      // ---- per-patch statistics (independent of the running state)
      float valid[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        float cnt = 0.f, sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          valid[j][k] = ((mk[j] >> (8 * k)) & 0xffu) ? 0.f : 1.f;
          cnt += valid[j][k];
          sum += xv[k] * valid[j][k];
        }
        cnt = group8_sum(cnt);
        sum = group8_sum(sum);
        const float cnt_safe = cnt == 0.f ? 1.f : cnt;
        const float inc_mu = cnt == 0.f ? 0.f : __fdiv_rn(sum, cnt_safe);
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = (xv[k] - inc_mu) * valid[j][k];
          sq += d * d;
        }
        sq = group8_sum(sq);
        const float inc_var = cnt == 0.f ? 0.f : __fdiv_rn(sq, cnt_safe);
        const float inc_sigma = sqrtf(fmaxf(inc_var, 0.f));
        const int pi = grp + 4 * j;  // patch index inside the chunk
        if (q == 0) {
          s_inc[warp][pi][0] = cnt;
          s_inc[warp][pi][1] = inc_mu;
          s_inc[warp][pi][2] = inc_sigma;
        }
        // the patch counts as padded iff its LAST element is padded (timesfm.py:97)
        const bool last_padded = (mk[j] >> 24) != 0;
        const int patch = (c0 >> 5) + pi;
        if (q == 7 && patch < num_patches) {
          if (patch_mask_out != nullptr) patch_mask_out[b * num_patches + patch] = last_padded ? 1 : 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, q == 7 && patch < num_patches && last_padded);
        masked_patches += __popc(bal);
      }
      __syncwarp();

      // ---- sequential merge over the chunk's patches, in the reference's order and formula
      // (update_running_stats; HF twin modeling_timesfm2_5.py:528-568).  Every lane runs the scan
      // and keeps the (mu, sigma) of the patches it owns.
      float my_mu[4], my_sigma[4];
      float keep_mu = 0.f, keep_sigma = 0.f;
      const int chunk_patches = min(16, num_patches - (c0 >> 5));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < chunk_patches) {
          const float inc_n = s_inc[warp][i][0], inc_mu = s_inc[warp][i][1], inc_sigma = s_inc[warp][i][2];
          const float new_n = __fadd_rn(run_n, inc_n);
          const float new_n_safe = new_n == 0.f ? 1.f : new_n;
          float new_mu = __fdiv_rn(__fadd_rn(__fmul_rn(run_n, run_mu), __fmul_rn(inc_mu, inc_n)), new_n_safe);
          if (new_n == 0.f) new_mu = 0.f;
          const float d1 = __fsub_rn(run_mu, new_mu), d2 = __fsub_rn(inc_mu, new_mu);
          const float t1 = __fmul_rn(run_n, __fmul_rn(run_sigma, run_sigma));
          const float t2 = __fmul_rn(inc_n, __fmul_rn(inc_sigma, inc_sigma));
          const float t3 = __fmul_rn(run_n, __fmul_rn(d1, d1));
          const float t4 = __fmul_rn(inc_n, __fmul_rn(d2, d2));
          float new_var = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4), new_n_safe);
          if (new_n == 0.f) new_var = 0.f;
          run_n = new_n;
          run_mu = new_mu;
          run_sigma = sqrtf(fmaxf(new_var, 0.f));
        }
        if ((i & 3) == grp) {
          my_mu[i >> 2] = run_mu;
          my_sigma[i >> 2] = run_sigma;
        }
        if (i == lane) {
          keep_mu = run_mu;
          keep_sigma = run_sigma;
        }
      }
      if (lane < chunk_patches) {
        if (mu_out != nullptr) mu_out[b * num_patches + (c0 >> 5) + lane] = keep_mu;
        if (sigma_out != nullptr) sigma_out[b * num_patches + (c0 >> 5) + lane] = keep_sigma;
      }
      __syncwarp();

      // ---- RevIN with the cumulative stats of the own patch, zero the padded points, emit tokens
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int patch = (c0 >> 5) + grp + 4 * j;
        if (patch < num_patches) {
          const float sig_safe = my_sigma[j] < 1e-6f ? 1.f : my_sigma[j];
          const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          float val[4], msk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            msk[k] = 1.f - valid[j][k];
            val[k] = valid[j][k] != 0.f ? __fdiv_rn(xv[k] - my_mu[j], sig_safe) : 0.f;
          }
          store_token_quad<OUT>(tokens, b * num_patches + patch, q, val, msk);
        }
      }
    }
    if (lane == 0 && num_masked_out != nullptr) num_masked_out[b] = masked_patches;
  }
}

// ----------------------------------------------------------------------------------------
// Chronos-2 context preparation.
// ----------------------------------------------------------------------------------------
template <int OUT>
__device__ __forceinline__ void store_elem(void* base, int64_t idx, int64_t lo_idx, float v) {
  if constexpr (OUT == TSFMX_DT_F32) {
    reinterpret_cast<float*>(base)[idx] = v;
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  } else {
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = h;
    reinterpret_cast<__nv_bfloat16*>(base)[lo_idx] = l;
  }
}

template <int OUT>
__global__ void __launch_bounds__(WARPS * 32) chronos2_patchify_norm_kernel(
    const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t batch, int context, int patch,
    int use_arcsinh, float time_scale, int out_cols, void* out, uint8_t* __restrict__ attn_mask,
    float* __restrict__ loc_out, float* __restrict__ scale_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_patches = (context + patch - 1) / patch;
  const int padded_len = num_patches * patch;
  const int left_pad = padded_len - context;

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    const uint8_t* mr = mask + b * context;
    // pass 1: nanmean
    float sum = 0.f, cnt = 0.f;
    for (int e = lane; e < context; e += 32) {
      const float v = __ldg(xr + e);
      if (!isnan(v)) { sum += v; cnt += 1.f; }
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    const float loc = cnt > 0.f ? __fdiv_rn(sum, cnt) : 0.f;  // nan_to_num(nanmean) -> 0
    // pass 2: sqrt(nanmean((x - loc)^2))
    float sq = 0.f;
    for (int e = lane; e < context; e += 32) {
      const float v = __ldg(xr + e);
      if (!isnan(v)) { const float d = v - loc; sq += d * d; }
    }
    sq = warp_sum(sq);
    float scale = cnt > 0.f ? sqrtf(__fdiv_rn(sq, cnt)) : 1.f;  // nan -> 1
    if (scale == 0.f) scale = 1e-5f;
    if (lane == 0) {
      if (loc_out != nullptr) loc_out[b] = loc;
      if (scale_out != nullptr) scale_out[b] = scale;
    }
    // pass 3: emit [time_enc | values | mask] per patch, rows of out_cols (tail zero-filled)
    const int row_elems = out_cols;
    for (int idx = lane; idx < num_patches * row_elems; idx += 32) {
      const int n = idx / row_elems, c = idx - n * row_elems;
      const int sect = c / patch, pp = c - sect * patch;
      float val = 0.f;
      if (sect < 3) {
        const int pe = n * patch + pp;     // index in the left-padded context
        const int src = pe - left_pad;     // index in the real context (< 0: NaN padding)
        if (sect == 0) {
          val = __fdiv_rn(static_cast<float>(pe - padded_len), time_scale);
        } else {
          const float m = (src >= 0 && mr[src] == 0) ? 1.f : 0.f;  // 1 = observed
          if (sect == 2) {
            val = m;
          } else if (m > 0.f) {
            float s = __fdiv_rn(xr[src] - loc, scale);
            if (use_arcsinh) s = asinhf(s);
            val = s;
          }
        }
      }
      const int64_t o = (b * num_patches + n) * static_cast<int64_t>(OUT == TSFMX_DT_BF16_SPLIT ? 2 * row_elems : row_elems);
      store_elem<OUT>(out, o + c, o + row_elems + c, val);
    }
    if (attn_mask != nullptr) {
      for (int n = lane; n < num_patches; n += 32) {
        int any = 0;
        for (int pp = 0; pp < patch; ++pp) {
          const int src = n * patch + pp - left_pad;
          any |= (src >= 0 && mr[src] == 0) ? 1 : 0;
        }
        attn_mask[b * num_patches + n] = static_cast<uint8_t>(any);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// Chronos-T5 mean-scale + uniform-bin tokeniser.  Bit-exact ids: fp32 IEEE division x / scale,
// then count(boundaries <= v) looked up in the caller's boundary table (torch.bucketize right=True).
// ----------------------------------------------------------------------------------------
constexpr int T5_MAX_BOUNDS = 8192;

__device__ __forceinline__ int bucketize_right(const float* __restrict__ sb, int nb, float v, float b1,
                                               float inv_step) {
  // number of boundaries <= v.  Uniform-grid guess (interior boundaries b[1..nb-2] are evenly
  // spaced), then an exact fix-up against the table, valid for any ascending table.
  float g = (v - b1) * inv_step;
  g = fminf(fmaxf(g, -1.f), static_cast<float>(nb));
  int i = static_cast<int>(floorf(g)) + 2;
  i = max(0, min(nb, i));
  while (i > 0 && !(sb[i - 1] <= v)) --i;
  while (i < nb && sb[i] <= v) ++i;
  return i;
}

template <int NV>  // float4 per lane cached in registers (context <= 128 * NV); NV == 0: re-read
__global__ void __launch_bounds__(WARPS * 32) chronos_t5_tokenize_kernel(
    const float* __restrict__ x, int64_t batch, int context, const float* __restrict__ boundaries, int nb,
    int n_special, int n_tokens, int pad_id, int eos_id, int64_t* __restrict__ ids,
    uint8_t* __restrict__ attn_mask, float* __restrict__ scale_out) {
  extern __shared__ float sb[];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sb[i] = boundaries[i];
  __syncthreads();
  const float b1 = nb > 2 ? sb[1] : 0.f;
  const float inv_step = nb > 3 ? static_cast<float>(nb - 3) / (sb[nb - 2] - sb[1]) : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (context % 4) == 0;
  const int nvec = context >> 2;

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    int64_t* idr = ids + b * (context + 1);
    uint8_t* amr = attn_mask + b * (context + 1);
    float4 cache[NV > 0 ? NV : 1];
    // sum(|x|) is accumulated in fp64 and rounded to fp32 once, so the scale (and hence every id)
    // does not depend on the reduction order.
    double sum = 0.0;
    float cnt = 0.f;
    if (NV > 0 && vec) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int f = lane + 32 * j;
        cache[j] = f < nvec ? ld_stream_f4(xr + 4 * f) : make_float4(NAN, NAN, NAN, NAN);
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float xv[4] = {cache[j].x, cache[j].y, cache[j].z, cache[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!isnan(xv[k])) { sum += static_cast<double>(fabsf(xv[k])); cnt += 1.f; }
      }
    } else {
      for (int e = lane; e < context; e += 32) {
        const float v = __ldg(xr + e);
        if (!isnan(v)) { sum += static_cast<double>(fabsf(v)); cnt += 1.f; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt = warp_sum(cnt);
    float scale = __fdiv_rn(static_cast<float>(sum), cnt);  // 0/0 -> NaN -> 1 below
    if (!(scale > 0.f)) scale = 1.f;
    if (lane == 0) {
      if (scale_out != nullptr) scale_out[b] = scale;
      idr[context] = eos_id;
      amr[context] = 1;
    }
    auto tok = [&](float v) -> int64_t {
      if (isnan(v)) return pad_id;
      int t = bucketize_right(sb, nb, __fdiv_rn(v, scale), b1, inv_step) + n_special;
      t = max(0, min(n_tokens - 1, t));
      return t;
    };
    if (NV > 0 && vec) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int f = lane + 32 * j;
        if (f < nvec) {
          const float xv[4] = {cache[j].x, cache[j].y, cache[j].z, cache[j].w};
          uint32_t am = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            idr[4 * f + k] = tok(xv[k]);
            am |= (isnan(xv[k]) ? 0u : 1u) << (8 * k);
          }
          // row base is only byte-aligned (row length C+1): byte stores
#pragma unroll
          for (int k = 0; k < 4; ++k) amr[4 * f + k] = static_cast<uint8_t>((am >> (8 * k)) & 0xffu);
        }
      }
    } else {
      for (int e = lane; e < context; e += 32) {
        const float v = __ldg(xr + e);
        idr[e] = tok(v);
        amr[e] = isnan(v) ? 0 : 1;
      }
    }
  }
}

__global__ void chronos_t5_dequantize_kernel(const int64_t* __restrict__ ids, int64_t total, int length,
                                             const float* __restrict__ centers, int n_centers, int n_special,
                                             const float* __restrict__ scale, float* __restrict__ values) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t idx = ids[i] - n_special - 1;
    idx = idx < 0 ? 0 : (idx > n_centers - 1 ? n_centers - 1 : idx);
    values[i] = __fmul_rn(__ldg(centers + idx), __ldg(scale + i / length));
  }
}

// fp32 rows -> bf16 / split bf16 rows
template <int OUT>
__global__ void cast_rows_kernel(const float* __restrict__ in, int64_t rows, int cols, int64_t ld_in, void* out) {
  const int vec_per_row = cols >> 2;
  const int64_t total = rows * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row) * 4;
    const float4 v = ld_stream_f4(in + r * ld_in + c);
    if constexpr (OUT == TSFMX_DT_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + r * cols + c;
      st_stream_u2(o, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + r * 2 * cols + c;
      uint2 h, l;
      split_bf16x2(v.x, v.y, h.x, l.x);
      split_bf16x2(v.z, v.w, h.y, l.y);
      st_stream_u2(o, h);
      st_stream_u2(o + cols, l);
    }
  }
}

int grid_for_series(int64_t batch) {
  const int64_t blocks = (batch + WARPS - 1) / WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8 * 4;  // a few waves of a persistent-style grid
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_timesfm_patchify_norm(const float* x, const uint8_t* mask, int64_t batch, int32_t context,
                                           int32_t patch_len, int32_t tokens_dtype, void* tokens, float* mu,
                                           float* sigma, uint8_t* patch_mask, int32_t* num_masked, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(x != nullptr && mask != nullptr && tokens != nullptr, "timesfm_patchify_norm: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && context > 0, "timesfm_patchify_norm: bad sizes");
  if (patch_len != 32) {
    set_error("timesfm_patchify_norm: patch_len %d unsupported (TimesFM 2.5 uses 32)", patch_len);
    return TSFMX_ERR_UNSUPPORTED;
  }
  TSFMX_REQUIRE(context % patch_len == 0, "context length (%d) must be divisible by patch length (%d)", context,
                patch_len);
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(mask) % 4 == 0 &&
                    reinterpret_cast<uintptr_t>(tokens) % 16 == 0,
                "timesfm_patchify_norm: x/tokens must be 16-byte and mask 4-byte aligned");
  if (batch == 0) return TSFMX_OK;
  const int grid = grid_for_series(batch);
  const dim3 block(WARPS * 32);
  switch (tokens_dtype) {
    case TSFMX_DT_F32:
      timesfm_patchify_norm_kernel<TSFMX_DT_F32><<<grid, block, 0, stream>>>(x, mask, batch, context, tokens, mu, sigma,
                                                                            patch_mask, num_masked);
      break;
    case TSFMX_DT_BF16:
      timesfm_patchify_norm_kernel<TSFMX_DT_BF16><<<grid, block, 0, stream>>>(x, mask, batch, context, tokens, mu,
                                                                             sigma, patch_mask, num_masked);
      break;
    case TSFMX_DT_BF16_SPLIT:
      timesfm_patchify_norm_kernel<TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(x, mask, batch, context, tokens, mu,
                                                                                   sigma, patch_mask, num_masked);
      break;
    default:
      set_error("timesfm_patchify_norm: bad tokens_dtype %d", tokens_dtype);
      return TSFMX_ERR_INVALID_ARGUMENT;
  }
  return check_last_launch("timesfm_patchify_norm");
}

extern "C" int tsfmx_chronos2_patchify_norm(const float* x, const uint8_t* mask, int64_t batch, int32_t context,
                                            int32_t patch, int32_t use_arcsinh, float time_encoding_scale,
                                            int32_t out_dtype, int32_t out_cols, void* patched, uint8_t* attn_mask,
                                            float* loc, float* scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(x != nullptr && mask != nullptr && patched != nullptr, "chronos2_patchify_norm: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && context > 0 && patch > 0, "chronos2_patchify_norm: bad sizes");
  TSFMX_REQUIRE(out_cols >= 3 * patch, "chronos2_patchify_norm: out_cols (%d) < 3 * patch (%d)", out_cols, 3 * patch);
  if (batch == 0) return TSFMX_OK;
  const int grid = grid_for_series(batch);
  const dim3 block(WARPS * 32);
  switch (out_dtype) {
    case TSFMX_DT_F32:
      chronos2_patchify_norm_kernel<TSFMX_DT_F32><<<grid, block, 0, stream>>>(
          x, mask, batch, context, patch, use_arcsinh, time_encoding_scale, out_cols, patched, attn_mask, loc, scale);
      break;
    case TSFMX_DT_BF16:
      chronos2_patchify_norm_kernel<TSFMX_DT_BF16><<<grid, block, 0, stream>>>(
          x, mask, batch, context, patch, use_arcsinh, time_encoding_scale, out_cols, patched, attn_mask, loc, scale);
      break;
    case TSFMX_DT_BF16_SPLIT:
      chronos2_patchify_norm_kernel<TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(
          x, mask, batch, context, patch, use_arcsinh, time_encoding_scale, out_cols, patched, attn_mask, loc, scale);
      break;
    default:
      set_error("chronos2_patchify_norm: bad out_dtype %d", out_dtype);
      return TSFMX_ERR_INVALID_ARGUMENT;
  }
  return check_last_launch("chronos2_patchify_norm");
}

extern "C" int tsfmx_chronos_t5_tokenize(const float* x, int64_t batch, int32_t context, const float* boundaries,
                                         int32_t n_boundaries, int32_t n_special, int32_t n_tokens, int32_t pad_id,
                                         int32_t eos_id, int64_t* ids, uint8_t* attn_mask, float* scale,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(x != nullptr && boundaries != nullptr && ids != nullptr && attn_mask != nullptr,
                "chronos_t5_tokenize: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && context > 0, "chronos_t5_tokenize: bad sizes");
  TSFMX_REQUIRE(n_boundaries >= 1 && n_boundaries <= T5_MAX_BOUNDS, "chronos_t5_tokenize: n_boundaries out of range");
  if (batch == 0) return TSFMX_OK;
  const int grid = grid_for_series(batch);
  const dim3 block(WARPS * 32);
  const size_t smem = static_cast<size_t>(n_boundaries) * sizeof(float);
  const bool aligned = reinterpret_cast<uintptr_t>(x) % 16 == 0 && context % 4 == 0;
  if (aligned && context <= 512) {
    chronos_t5_tokenize_kernel<4><<<grid, block, smem, stream>>>(x, batch, context, boundaries, n_boundaries, n_special,
                                                                 n_tokens, pad_id, eos_id, ids, attn_mask, scale);
  } else if (aligned && context <= 2048) {
    chronos_t5_tokenize_kernel<16><<<grid, block, smem, stream>>>(x, batch, context, boundaries, n_boundaries,
                                                                  n_special, n_tokens, pad_id, eos_id, ids, attn_mask,
                                                                  scale);
  } else {
    chronos_t5_tokenize_kernel<0><<<grid, block, smem, stream>>>(x, batch, context, boundaries, n_boundaries, n_special,
                                                                 n_tokens, pad_id, eos_id, ids, attn_mask, scale);
  }
  return check_last_launch("chronos_t5_tokenize");
}

extern "C" int tsfmx_chronos_t5_dequantize(const int64_t* ids, int64_t batch, int32_t length, const float* centers,
                                           int32_t n_centers, int32_t n_special, const float* scale, float* values,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(ids != nullptr && centers != nullptr && scale != nullptr && values != nullptr,
                "chronos_t5_dequantize: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && length > 0 && n_centers > 0, "chronos_t5_dequantize: bad sizes");
  const int64_t total = batch * length;
  if (total == 0) return TSFMX_OK;
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
  chronos_t5_dequantize_kernel<<<grid, 256, 0, stream>>>(ids, total, length, centers, n_centers, n_special, scale,
                                                         values);
  return check_last_launch("chronos_t5_dequantize");
}

extern "C" int tsfmx_cast_rows(const float* in, int64_t rows, int32_t cols, int64_t ld_in, int32_t out_dtype,
                               void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(in != nullptr && out != nullptr, "cast_rows: NULL pointer");
  TSFMX_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && ld_in % 4 == 0 && ld_in >= cols,
                "cast_rows: cols and ld_in must be multiples of 4");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0,
                "cast_rows: misaligned pointer");
  if (rows == 0) return TSFMX_OK;
  const int64_t total = rows * (cols / 4);
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
  if (out_dtype == TSFMX_DT_BF16) {
    cast_rows_kernel<TSFMX_DT_BF16><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, out);
  } else if (out_dtype == TSFMX_DT_BF16_SPLIT) {
    cast_rows_kernel<TSFMX_DT_BF16_SPLIT><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, out);
  } else {
    set_error("cast_rows: bad out_dtype %d", out_dtype);
    return TSFMX_ERR_INVALID_ARGUMENT;
  }
  return check_last_launch("cast_rows");
}

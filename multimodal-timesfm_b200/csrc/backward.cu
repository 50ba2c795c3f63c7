// Backward-pass kernels of the fusion fine-tune step (reference tsfmx/trainer.py:200-219: the adapter is
// frozen, trainer.py:76-77, so only activation gradients flow through the backbone and weight gradients
// are produced for the fusion Linear layers alone).
//
//   rmsnorm_bwd_chain       g_total = g_res + RMSNorm_bwd(v1, w1, g1);  g2 = RMSNorm_bwd(v2, w2, g_total)
//                           — the two norm/residual junctions of a TimesFM 2.5 layer, one pass each
//   timesfm_attention_bwd   d(q, k, v) of the attention core incl. RoPE / q-k RMSNorm / per-dim scale
//   transpose_mask          [R, C] -> [C, Rpad] (optionally gated by relu'(mask)) so the fusion wgrad
//                           dW = dpre^T . text is a K-major tcgen05 GEMM with K = tokens
//   mask_cast_rows          g * relu'(mask) -> bf16 / split rows (dgrad operand of multi-layer fusion)
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ float4 load4(const void* base, int dtype, int64_t idx) {
  if (dtype == TSFMX_DT_F32) return ld_stream_f4(reinterpret_cast<const float*>(base) + idx);
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
  const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1));
}

template <int OUT>
__device__ __forceinline__ void store4(void* out, int64_t row, int cols, int c, float4 v) {
  if constexpr (OUT == TSFMX_DT_F32) {
    st_stream_f4(reinterpret_cast<float*>(out) + row * cols + c, v);
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    st_stream_u2(reinterpret_cast<__nv_bfloat16*>(out) + row * cols + c,
                 make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * 2 * cols + c;
    uint2 h, l;
    split_bf16x2(v.x, v.y, h.x, l.x);
    split_bf16x2(v.z, v.w, h.y, l.y);
    st_stream_u2(o, h);
    st_stream_u2(o + cols, l);
  }
}

// dv = r * (w*g) - v * r^3 * mean(v * w * g),  r = (mean(v^2) + eps)^-1/2     (in place on g)
// acc (full fine-tune): this warp's partial of the scale gradient, acc[c] += g[c] * v[c] * r, in shared memory.
template <int NV>
__device__ __forceinline__ void rms_bwd_row(const float4 (&v)[NV], float4 (&g)[NV], const float* __restrict__ w,
                                            int lane, float eps, float* acc = nullptr) {
  constexpr int COLS = NV * 128;
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  ss = warp_sum(ss);
  const float r = 1.0f / sqrtf(ss / static_cast<float>(COLS) + eps);
  if (acc != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4* ap = reinterpret_cast<float4*>(acc + 4 * (lane + 32 * j));
      float4 a = *ap;
      a.x = fmaf(g[j].x, v[j].x * r, a.x), a.y = fmaf(g[j].y, v[j].y * r, a.y);
      a.z = fmaf(g[j].z, v[j].z * r, a.z), a.w = fmaf(g[j].w, v[j].w * r, a.w);
      *ap = a;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + 4 * (lane + 32 * j)));
    g[j].x *= ww.x, g[j].y *= ww.y, g[j].z *= ww.z, g[j].w *= ww.w;
    dot += v[j].x * g[j].x + v[j].y * g[j].y + v[j].z * g[j].z + v[j].w * g[j].w;
  }
  dot = warp_sum(dot);
  const float c = r * r * r * dot / static_cast<float>(COLS);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    g[j].x = r * g[j].x - v[j].x * c, g[j].y = r * g[j].y - v[j].y * c;
    g[j].z = r * g[j].z - v[j].z * c, g[j].w = r * g[j].w - v[j].w * c;
  }
}

// WG (full fine-tune): the gradients of the two norm scales leave the same pass - dw1[c] += sum_r g1 * v1_hat,
// dw2[c] += sum_r g_total * v2_hat - through warp-private column partials in shared memory, folded per block and added
// to the pre-zeroed outputs (the separate colsum_wgrad launches re-read g and v: 4 more passes per layer).
template <int NV, int OUT, bool WG = false>
__global__ void __launch_bounds__(WARPS * 32, 2) rmsnorm_bwd_chain_kernel(
    const float* g_res, const void* __restrict__ v1, int v1_dtype, const float* __restrict__ w1,
    const void* __restrict__ g1, int g1_dtype, const void* __restrict__ v2, int v2_dtype,
    const float* __restrict__ w2, int64_t rows, float eps, float* g_total, void* g2, float* __restrict__ dw1 = nullptr,
    float* __restrict__ dw2 = nullptr) {
  constexpr int COLS = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ __align__(16) float s_wg[];  // WG: [2][WARPS][COLS]
  float* acc1 = nullptr;
  float* acc2 = nullptr;
  if constexpr (WG) {
    for (int i = threadIdx.x; i < 2 * WARPS * COLS; i += blockDim.x) s_wg[i] = 0.f;
    __syncthreads();
    if (dw1 != nullptr) acc1 = s_wg + warp * COLS;
    if (dw2 != nullptr) acc2 = s_wg + (WARPS + warp) * COLS;
  }
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * WARPS + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * WARPS) {
    float4 g[NV], v[NV];
    if (v1 != nullptr) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t idx = r * COLS + 4 * (lane + 32 * j);
        v[j] = load4(v1, v1_dtype, idx);
        g[j] = load4(g1, g1_dtype, idx);
      }
      rms_bwd_row<NV>(v, g, w1, lane, eps, acc1);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (g_res != nullptr) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 a = ld_stream_f4(g_res + r * COLS + 4 * (lane + 32 * j));
        g[j].x += a.x, g[j].y += a.y, g[j].z += a.z, g[j].w += a.w;
      }
    }
    if (g_total != nullptr) {
#pragma unroll
      for (int j = 0; j < NV; ++j) st_stream_f4(g_total + r * COLS + 4 * (lane + 32 * j), g[j]);
    }
    if (v2 != nullptr) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = load4(v2, v2_dtype, r * COLS + 4 * (lane + 32 * j));
      rms_bwd_row<NV>(v, g, w2, lane, eps, acc2);
#pragma unroll
      for (int j = 0; j < NV; ++j) store4<OUT>(g2, r, COLS, 4 * (lane + 32 * j), g[j]);
    }
  }
  if constexpr (WG) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * COLS; i += blockDim.x) {
      const int which = i / COLS, c = i - which * COLS;
      float* dst = which == 0 ? dw1 : dw2;
      if (dst == nullptr) continue;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += s_wg[(which * WARPS + w) * COLS + c];
      atomicAdd(dst + c, t);
    }
  }
}

// ----------------------------------------------------------------------------------------
// Column reductions of the full fine-tune ("baseline" mode, reference tsfmx/trainer.py:78-79): gradient of an RMSNorm
// scale, dscale[c] = sum_r g[r, c] * v[r, c] * rsqrt(mean(v[r]^2) + eps), or (normalize = 0) a bias gradient
// dbias[c] = sum_r g[r, c].  Rows stream once; every warp keeps its column partials in registers, the block folds
// them in shared memory and adds one value per column to the (pre-zeroed) output.
// ----------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(WARPS * 32) colsum_wgrad_kernel(const void* __restrict__ v, int v_dtype,
                                                                   const void* __restrict__ g, int g_dtype, int64_t rows,
                                                                   float eps, int normalize, float* __restrict__ out) {
  constexpr int COLS = NV * 128;
  __shared__ float s_part[WARPS][COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * WARPS + warp; r < rows; r += static_cast<int64_t>(gridDim.x) * WARPS) {
    float4 gg[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) gg[j] = load4(g, g_dtype, r * COLS + 4 * (lane + 32 * j));
    if (normalize) {
      float4 vv[NV];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        vv[j] = load4(v, v_dtype, r * COLS + 4 * (lane + 32 * j));
        ss += vv[j].x * vv[j].x + vv[j].y * vv[j].y + vv[j].z * vv[j].z + vv[j].w * vv[j].w;
      }
      ss = warp_sum(ss);
      const float rs = 1.0f / sqrtf(ss / static_cast<float>(COLS) + eps);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        acc[j].x = fmaf(gg[j].x, vv[j].x * rs, acc[j].x), acc[j].y = fmaf(gg[j].y, vv[j].y * rs, acc[j].y);
        acc[j].z = fmaf(gg[j].z, vv[j].z * rs, acc[j].z), acc[j].w = fmaf(gg[j].w, vv[j].w * rs, acc[j].w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[j].x += gg[j].x, acc[j].y += gg[j].y, acc[j].z += gg[j].z, acc[j].w += gg[j].w;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) *reinterpret_cast<float4*>(&s_part[warp][4 * (lane + 32 * j)]) = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < COLS; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) t += s_part[w][c];
    atomicAdd(out + c, t);
  }
}

// Bias gradient for any width that is a multiple of 4 (the 3072-wide hidden layers and the 336 -> 384 wide output
// block of Chronos-2): dbias[c] = sum_r g[r, c].  A block owns 128 columns (32 threads x float4) and a slice of the
// rows (8 row lanes); partials meet in shared memory and leave as one atomicAdd per column.
__global__ void __launch_bounds__(256) colsum_bias_kernel(const void* __restrict__ g, int g_dtype, int64_t rows, int cols,
                                                          int64_t rows_per_block, float* __restrict__ out) {
  __shared__ float4 s_part[8][32];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128 + 4 * cq;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c0 < cols) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float4 v = load4(g, g_dtype, r * cols + c0);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  s_part[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && c0 < cols) {
    float4 t = s_part[0][cq];
#pragma unroll
    for (int w = 1; w < 8; ++w) t.x += s_part[w][cq].x, t.y += s_part[w][cq].y, t.z += s_part[w][cq].z, t.w += s_part[w][cq].w;
    atomicAdd(out + c0, t.x), atomicAdd(out + c0 + 1, t.y), atomicAdd(out + c0 + 2, t.z), atomicAdd(out + c0 + 3, t.w);
  }
}

// ----------------------------------------------------------------------------------------
// attention backward (fp32 SIMT; one warp per (series, head))
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float load_elem(const void* base, int dtype, int64_t idx) {
  return dtype == TSFMX_DT_F32 ? reinterpret_cast<const float*>(base)[idx]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

template <int HD, int OUT>
__global__ void timesfm_attention_bwd_kernel(const void* __restrict__ qkv, int qkv_dtype, const void* __restrict__ dout,
                                             int dout_dtype, int64_t batch, int num_patches, int num_heads,
                                             const uint8_t* __restrict__ patch_mask,
                                             const int32_t* __restrict__ num_masked,
                                             const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w,
                                             const float* __restrict__ k_ln_w, const float* __restrict__ q_scale,
                                             float eps, void* dqkv, float* __restrict__ dparams) {
  constexpr int DPL = (HD + 31) / 32;
  // full fine-tune only (dparams != NULL): d/d(q_ln_w * q_scale)[d] = sum dq'[d] * qhat[d], d/d(k_ln_w)[d] = sum dk'[d] * khat[d]
  float dwq[DPL], dwk[DPL];
#pragma unroll
  for (int t = 0; t < DPL; ++t) dwq[t] = 0.f, dwk[t] = 0.f;
  constexpr int HALF = HD / 2;
  constexpr int LD = HD + 1;
  extern __shared__ float smem[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = num_patches;
  const int per_warp = 9 * N * LD + 4 * N;
  float* sQr = smem + warp * per_warp;  // RoPE'd q (pre-norm)
  float* sKr = sQr + N * LD;
  float* sQ = sKr + N * LD;             // conditioned q' = wq * qr * r_q
  float* sK = sQ + N * LD;
  float* sV = sK + N * LD;
  float* sDO = sV + N * LD;
  float* sDQ = sDO + N * LD;            // dL/dq'
  float* sDK = sDQ + N * LD;
  float* sDV = sDK + N * LD;
  float* sRq = sDV + N * LD;            // [N] row rsqrt of q
  float* sRk = sRq + N;
  float* sP = sRk + N;
  float* sDS = sP + N;
  const int width = num_heads * HD;
  const int64_t qkv_ld = 3 * static_cast<int64_t>(width);
  const int64_t total = batch * num_heads;

  for (int64_t wu = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; wu < total;
       wu += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t b = wu / num_heads;
    const int h = static_cast<int>(wu - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;
    const uint8_t* pm = patch_mask != nullptr ? patch_mask + b * N : nullptr;

    // ---- recompute the forward conditioning
    for (int n = 0; n < N; ++n) {
      const int64_t base = (b * N + n) * qkv_ld + h * HD;
      const int64_t obase = (b * N + n) * static_cast<int64_t>(width) + h * HD;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          sQ[n * LD + d] = load_elem(qkv, qkv_dtype, base + d);             // raw, rotated below
          sK[n * LD + d] = load_elem(qkv, qkv_dtype, base + width + d);
          sV[n * LD + d] = load_elem(qkv, qkv_dtype, base + 2 * width + d);
          sDO[n * LD + d] = load_elem(dout, dout_dtype, obase + d);
          sDQ[n * LD + d] = 0.f, sDK[n * LD + d] = 0.f, sDV[n * LD + d] = 0.f;
        }
      }
    }
    __syncwarp();
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
          float sn, cs;
          sincosf(pos * __ldg(inv_freq + f), &sn, &cs);
          const int dp = d < HALF ? d + HALF : d - HALF;
          const float sgn = d < HALF ? -1.f : 1.f;
          qr[t] = sQ[n * LD + d] * cs + sgn * sQ[n * LD + dp] * sn;
          kr[t] = sK[n * LD + d] * cs + sgn * sK[n * LD + dp] * sn;
          qss += qr[t] * qr[t];
          kss += kr[t] * kr[t];
        }
      }
      qss = warp_sum(qss);
      kss = warp_sum(kss);
      const float qrs = 1.0f / sqrtf(qss / static_cast<float>(HD) + eps);
      const float krs = 1.0f / sqrtf(kss / static_cast<float>(HD) + eps);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          sQr[n * LD + d] = qr[t];
          sKr[n * LD + d] = kr[t];
          sQ[n * LD + d] = __ldg(q_ln_w + d) * __ldg(q_scale + d) * (qr[t] * qrs);
          sK[n * LD + d] = __ldg(k_ln_w + d) * (kr[t] * krs);
        }
      }
      if (lane == 0) sRq[n] = qrs, sRk[n] = krs;
    }
    __syncwarp();

    // ---- per query row: P, dP, dS, and the accumulations into dq', dk', dv
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
            for (int d = 0; d < HD; ++d) acc = fmaf(sQ[i * LD + d], sK[j * LD + d], acc);
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
          const float p = row_has_key ? (sP[j] == -INFINITY ? 0.f : expf(sP[j] - mx)) : 1.f;
          sP[j] = p;
          sum += p;
        }
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      __syncwarp();
      const int jend = row_has_key ? i + 1 : N;
      // dP_j = dO_i . v_j ; delta = sum_j P_j dP_j ; dS_j = P_j (dP_j - delta).  An all-masked row has uniform P
      // over all N keys and, like the reference's additive mask under autograd, still passes dS to q and k.
      float delta = 0.f;
      for (int j0 = 0; j0 < jend; j0 += 32) {
        const int j = j0 + lane;
        if (j < jend) {
          float acc = 0.f;
#pragma unroll 8
          for (int d = 0; d < HD; ++d) acc = fmaf(sDO[i * LD + d], sV[j * LD + d], acc);
          const float p = sP[j] * inv;
          sP[j] = p;
          sDS[j] = acc;
          delta += p * acc;
        }
      }
      delta = warp_sum(delta);
      __syncwarp();
      for (int j0 = 0; j0 < jend; j0 += 32) {
        const int j = j0 + lane;
        if (j < jend) sDS[j] = sP[j] * (sDS[j] - delta);
      }
      __syncwarp();
      float dq[DPL];
#pragma unroll
      for (int t = 0; t < DPL; ++t) dq[t] = 0.f;
      for (int j = 0; j < jend; ++j) {
        const float p = sP[j], ds = sDS[j];
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
          const int d = lane + 32 * t;
          if (d < HD) {
            sDV[j * LD + d] = fmaf(p, sDO[i * LD + d], sDV[j * LD + d]);
            sDK[j * LD + d] = fmaf(ds, sQ[i * LD + d], sDK[j * LD + d]);
            dq[t] = fmaf(ds, sK[j * LD + d], dq[t]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) sDQ[i * LD + d] = dq[t];
      }
      __syncwarp();
    }

    // ---- back through the conditioning: per-dim scale * RMSNorm, then the inverse rotation
    for (int n = 0; n < N; ++n) {
      const float pos = static_cast<float>(n - nm);
      const float rq = sRq[n], rk = sRk[n];
      float tq[DPL], tk[DPL];
      float cq = 0.f, ck = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        tq[t] = 0.f, tk[t] = 0.f;
        if (d < HD) {
          tq[t] = __ldg(q_ln_w + d) * __ldg(q_scale + d) * sDQ[n * LD + d];
          tk[t] = __ldg(k_ln_w + d) * sDK[n * LD + d];
          dwq[t] = fmaf(sDQ[n * LD + d], sQr[n * LD + d] * rq, dwq[t]);
          dwk[t] = fmaf(sDK[n * LD + d], sKr[n * LD + d] * rk, dwk[t]);
          cq += sQr[n * LD + d] * tq[t];
          ck += sKr[n * LD + d] * tk[t];
        }
      }
      cq = warp_sum(cq) / static_cast<float>(HD);
      ck = warp_sum(ck) / static_cast<float>(HD);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          sDQ[n * LD + d] = rq * tq[t] - sQr[n * LD + d] * rq * rq * rq * cq;  // d/d rope(q)
          sDK[n * LD + d] = rk * tk[t] - sKr[n * LD + d] * rk * rk * rk * ck;
        }
      }
      __syncwarp();
      const int64_t base = (b * N + n) * qkv_ld + h * HD;
#pragma unroll
      for (int t = 0; t < DPL; ++t) {
        const int d = lane + 32 * t;
        if (d < HD) {
          const int f = d < HALF ? d : d - HALF;
          float sn, cs;
          sincosf(pos * __ldg(inv_freq + f), &sn, &cs);
          const int dp = d < HALF ? d + HALF : d - HALF;
          const float sgn = d < HALF ? 1.f : -1.f;  // transpose of the forward rotation
          const float gq = sDQ[n * LD + d] * cs + sgn * sDQ[n * LD + dp] * sn;
          const float gk = sDK[n * LD + d] * cs + sgn * sDK[n * LD + dp] * sn;
          const float gv = sDV[n * LD + d];
          if constexpr (OUT == TSFMX_DT_F32) {
            float* o = reinterpret_cast<float*>(dqkv);
            o[base + d] = gq, o[base + width + d] = gk, o[base + 2 * width + d] = gv;
          } else if constexpr (OUT == TSFMX_DT_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dqkv);
            o[base + d] = __float2bfloat16_rn(gq);
            o[base + width + d] = __float2bfloat16_rn(gk);
            o[base + 2 * width + d] = __float2bfloat16_rn(gv);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dqkv);
            const int64_t row = (b * N + n) * 2 * qkv_ld;  // split rows: [hi(3W) | lo(3W)]
            const int64_t c = h * HD + d;
            split_bf16(gq, o[row + c], o[row + qkv_ld + c]);
            split_bf16(gk, o[row + width + c], o[row + qkv_ld + width + c]);
            split_bf16(gv, o[row + 2 * width + c], o[row + qkv_ld + 2 * width + c]);
          }
        }
      }
      __syncwarp();
    }
  }
  if (dparams != nullptr) {
#pragma unroll
    for (int t = 0; t < DPL; ++t) {
      const int d = lane + 32 * t;
      if (d < HD) {
        atomicAdd(dparams + d, dwq[t]);
        atomicAdd(dparams + HD + d, dwk[t]);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// transpose (+ relu' gate) and gated cast
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float load_any(const void* base, int dtype, int64_t row, int64_t ld, int col, int cols) {
  if (dtype == TSFMX_DT_F32) return reinterpret_cast<const float*>(base)[row * ld + col];
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + row * ld;
  if (dtype == TSFMX_DT_BF16) return __bfloat162float(p[col]);
  return __bfloat162float(p[col]) + __bfloat162float(p[cols + col]);  // split: hi + lo
}

template <int OUT>
__global__ void transpose_mask_kernel(const void* __restrict__ in, int in_dtype, int64_t rows, int cols, int64_t ld_in,
                                      const void* __restrict__ mask, int mask_dtype, int64_t ld_mask, void* out,
                                      int64_t ld_out) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = load_any(in, in_dtype, r, ld_in, c, cols);
      if (mask != nullptr && !(load_any(mask, mask_dtype, r, ld_mask, c, cols) > 0.f)) v = 0.f;
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < cols && r < ld_out) {
      const float v = tile[threadIdx.x][i];  // zero for the padding columns r >= rows
      if constexpr (OUT == TSFMX_DT_BF16) {
        reinterpret_cast<__nv_bfloat16*>(out)[c * ld_out + r] = __float2bfloat16_rn(v);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + c * 2 * ld_out;
        split_bf16(v, o[r], o[ld_out + r]);
      }
    }
  }
}

// 64 x 64 tiles, two elements per thread in both directions: every warp instruction moves 128 contiguous bytes of bf16
// (the 32 x 32 scalar version above moved 64 and ran at a quarter of the HBM peak; it stays as the fallback for odd
// sizes).  Needs cols and ld_out even, bf16 / fp32 rows whose pairs are naturally aligned.
__device__ __forceinline__ float2 load_pair(const void* base, int dtype, int64_t row, int64_t ld, int col, int cols) {
  if (dtype == TSFMX_DT_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + row * ld + col);
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + row * ld;
  const uint32_t h = *reinterpret_cast<const uint32_t*>(p + col);
  float2 v = make_float2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
  if (dtype == TSFMX_DT_BF16_SPLIT) {
    const uint32_t l = *reinterpret_cast<const uint32_t*>(p + cols + col);
    v.x += __uint_as_float(l << 16), v.y += __uint_as_float(l & 0xffff0000u);
  }
  return v;
}

template <int OUT>
__global__ void __launch_bounds__(256) transpose_mask64_kernel(const void* __restrict__ in, int in_dtype, int64_t rows,
                                                               int cols, int64_t ld_in, const void* __restrict__ mask,
                                                               int mask_dtype, int64_t ld_mask, void* out,
                                                               int64_t ld_out) {
  __shared__ float tile[64][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int c = c0 + 2 * tx;
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    const int64_t r = r0 + i;
    float2 v = make_float2(0.f, 0.f);
    if (r < rows && c < cols) {
      v = load_pair(in, in_dtype, r, ld_in, c, cols);
      if (mask != nullptr) {
        const float2 m = load_pair(mask, mask_dtype, r, ld_mask, c, cols);
        v.x = m.x > 0.f ? v.x : 0.f, v.y = m.y > 0.f ? v.y : 0.f;
      }
    }
    tile[i][2 * tx] = v.x;
    tile[i][2 * tx + 1] = v.y;
  }
  __syncthreads();
  const int64_t r = r0 + 2 * tx;  // output column pair (zero for the padding columns r >= rows)
#pragma unroll
  for (int j = ty; j < 64; j += 8) {
    const int oc = c0 + j;
    if (oc < cols && r < ld_out) {
      const float a = tile[2 * tx][j], b = tile[2 * tx + 1][j];
      if constexpr (OUT == TSFMX_DT_BF16) {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + oc * ld_out + r) = pack_bf16x2(a, b);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + oc * 2 * ld_out;
        uint32_t h, l;
        split_bf16x2(a, b, h, l);
        *reinterpret_cast<uint32_t*>(o + r) = h;
        *reinterpret_cast<uint32_t*>(o + ld_out + r) = l;
      }
    }
  }
}

template <int OUT>
__global__ void mask_cast_rows_kernel(const float* __restrict__ in, int64_t rows, int cols, const void* __restrict__ mask,
                                      int mask_dtype, int64_t ld_mask, void* out) {
  const int vec_per_row = cols >> 2;
  const int64_t total = rows * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row) * 4;
    float4 v = ld_stream_f4(in + r * cols + c);
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = load_any(mask, mask_dtype, r, ld_mask, c + k, cols);
    v.x = m[0] > 0.f ? v.x : 0.f, v.y = m[1] > 0.f ? v.y : 0.f;
    v.z = m[2] > 0.f ? v.z : 0.f, v.w = m[3] > 0.f ? v.w : 0.f;
    store4<OUT>(out, r, cols, c, v);
  }
}

int grid_for_rows(int64_t rows) {
  const int64_t blocks = (rows + WARPS - 1) / WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8 * 4;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

template <int NV>
int launch_bwd_chain(const float* g_res, const void* v1, int v1_dtype, const float* w1, const void* g1, int g1_dtype,
                     const void* v2, int v2_dtype, const float* w2, int64_t rows, float eps, float* g_total,
                     int g2_dtype, void* g2, float* dw1, float* dw2, cudaStream_t stream) {
  const dim3 block(WARPS * 32);
  if (dw1 != nullptr || dw2 != nullptr) {
    // scale gradients ride along: few, fat blocks (every block ends with one atomicAdd per column and output)
    constexpr int SMEM = 2 * WARPS * NV * 128 * 4;
    const int64_t blocks = (rows + WARPS - 1) / WARPS, cap = static_cast<int64_t>(num_sms()) * 2;
    const int grid = static_cast<int>(blocks < cap ? blocks : cap);
    auto go = [&](auto kern) -> int {
      static bool attr_set_dev[64] = {false};
      bool& attr_set = attr_set_dev[current_device()];
      if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
          set_error("rmsnorm_bwd_chain: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
          return TSFMX_ERR_CUDA;
        }
        attr_set = true;
      }
      kern<<<grid, block, SMEM, stream>>>(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2, v2_dtype, w2, rows, eps, g_total, g2,
                                          dw1, dw2);
      return check_last_launch("rmsnorm_bwd_chain");
    };
    if (g2_dtype == TSFMX_DT_F32) return go(rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_F32, true>);
    if (g2_dtype == TSFMX_DT_BF16) return go(rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_BF16, true>);
    return go(rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_BF16_SPLIT, true>);
  }
  const int grid = grid_for_rows(rows);
  if (g2_dtype == TSFMX_DT_F32)
    rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_F32><<<grid, block, 0, stream>>>(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2,
                                                                          v2_dtype, w2, rows, eps, g_total, g2);
  else if (g2_dtype == TSFMX_DT_BF16)
    rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_BF16><<<grid, block, 0, stream>>>(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2,
                                                                           v2_dtype, w2, rows, eps, g_total, g2);
  else
    rmsnorm_bwd_chain_kernel<NV, TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(g_res, v1, v1_dtype, w1, g1, g1_dtype,
                                                                                 v2, v2_dtype, w2, rows, eps, g_total, g2);
  return check_last_launch("rmsnorm_bwd_chain");
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_rmsnorm_bwd_chain_wgrad(const float* g_res, const void* v1, int32_t v1_dtype, const float* w1,
                                             const void* g1, int32_t g1_dtype, const void* v2, int32_t v2_dtype,
                                             const float* w2, int64_t rows, int32_t cols, float eps, float* g_total,
                                             int32_t g2_dtype, void* g2, float* dw1, float* dw2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(rows >= 0, "rmsnorm_bwd_chain: bad rows");
  TSFMX_REQUIRE(g_res != nullptr || v1 != nullptr, "rmsnorm_bwd_chain: no gradient input");
  TSFMX_REQUIRE(v1 == nullptr || (w1 != nullptr && g1 != nullptr), "rmsnorm_bwd_chain: v1 needs w1 and g1");
  TSFMX_REQUIRE(v2 == nullptr || (w2 != nullptr && g2 != nullptr), "rmsnorm_bwd_chain: v2 needs w2 and g2");
  TSFMX_REQUIRE(g_total != nullptr || g2 != nullptr, "rmsnorm_bwd_chain: no output requested");
  TSFMX_REQUIRE((dw1 == nullptr || v1 != nullptr) && (dw2 == nullptr || v2 != nullptr),
                "rmsnorm_bwd_chain: a scale gradient needs its norm (dw1 with v1, dw2 with v2)");
  auto dt_ok = [](int d) { return d == TSFMX_DT_F32 || d == TSFMX_DT_BF16; };
  TSFMX_REQUIRE((v1 == nullptr || (dt_ok(v1_dtype) && dt_ok(g1_dtype))) && (v2 == nullptr || dt_ok(v2_dtype)),
                "rmsnorm_bwd_chain: inputs must be f32 or bf16");
  TSFMX_REQUIRE(g2_dtype >= TSFMX_DT_F32 && g2_dtype <= TSFMX_DT_BF16_SPLIT, "rmsnorm_bwd_chain: bad g2_dtype");
  if (rows == 0) return TSFMX_OK;
  switch (cols) {
    case 1280:
      return launch_bwd_chain<10>(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2, v2_dtype, w2, rows, eps, g_total, g2_dtype,
                                  g2, dw1, dw2, stream);
    case 768:
      return launch_bwd_chain<6>(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2, v2_dtype, w2, rows, eps, g_total, g2_dtype,
                                 g2, dw1, dw2, stream);
    default:
      set_error("rmsnorm_bwd_chain: cols=%d unsupported (1280 or 768)", cols);
      return TSFMX_ERR_UNSUPPORTED;
  }
}

extern "C" int tsfmx_rmsnorm_bwd_chain(const float* g_res, const void* v1, int32_t v1_dtype, const float* w1,
                                       const void* g1, int32_t g1_dtype, const void* v2, int32_t v2_dtype,
                                       const float* w2, int64_t rows, int32_t cols, float eps, float* g_total,
                                       int32_t g2_dtype, void* g2, void* stream_) {
  return tsfmx_rmsnorm_bwd_chain_wgrad(g_res, v1, v1_dtype, w1, g1, g1_dtype, v2, v2_dtype, w2, rows, cols, eps, g_total,
                                       g2_dtype, g2, nullptr, nullptr, stream_);
}

extern "C" int tsfmx_colsum_wgrad(const void* v, int32_t v_dtype, const void* g, int32_t g_dtype, int64_t rows, int32_t cols,
                                  float eps, int32_t normalize, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(g != nullptr && out != nullptr && (!normalize || v != nullptr), "colsum_wgrad: NULL pointer");
  TSFMX_REQUIRE(rows >= 0, "colsum_wgrad: bad rows");
  auto dt_ok = [](int d) { return d == TSFMX_DT_F32 || d == TSFMX_DT_BF16; };
  TSFMX_REQUIRE(dt_ok(g_dtype) && (!normalize || dt_ok(v_dtype)), "colsum_wgrad: inputs must be f32 or bf16");
  if (rows == 0) return TSFMX_OK;
  int64_t blocks = (rows + WARPS - 1) / WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 2;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  switch (cols) {
    case 1280: colsum_wgrad_kernel<10><<<grid, WARPS * 32, 0, stream>>>(v, v_dtype, g, g_dtype, rows, eps, normalize, out); break;
    case 768: colsum_wgrad_kernel<6><<<grid, WARPS * 32, 0, stream>>>(v, v_dtype, g, g_dtype, rows, eps, normalize, out); break;
    default: {
      if (normalize || cols <= 0 || cols % 4 != 0) {
        set_error("colsum_wgrad: cols=%d unsupported (norm-scale gradients: 1280 or 768; bias gradients: any multiple of 4)",
                  cols);
        return TSFMX_ERR_UNSUPPORTED;
      }
      const int col_blocks = (cols + 127) / 128;
      int64_t row_blocks = (static_cast<int64_t>(num_sms()) * 4 + col_blocks - 1) / col_blocks;
      if (row_blocks > (rows + 63) / 64) row_blocks = (rows + 63) / 64;
      if (row_blocks < 1) row_blocks = 1;
      const int64_t rows_per_block = (rows + row_blocks - 1) / row_blocks;
      colsum_bias_kernel<<<dim3(col_blocks, static_cast<unsigned>(row_blocks)), 256, 0, stream>>>(g, g_dtype, rows, cols,
                                                                                                 rows_per_block, out);
    }
  }
  return check_last_launch("colsum_wgrad");
}

extern "C" int tsfmx_timesfm_attention_bwd(const void* qkv, int32_t qkv_dtype, const void* d_out, int32_t dout_dtype,
                                           int64_t batch, int32_t num_patches, int32_t num_heads, int32_t head_dim,
                                           const uint8_t* patch_mask, const int32_t* num_masked,
                                           const float* inv_freq, const float* q_ln_w, const float* k_ln_w,
                                           const float* q_scale, float eps, int32_t dqkv_dtype, void* dqkv,
                                           float* dparams, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && d_out != nullptr && dqkv != nullptr && inv_freq != nullptr && q_ln_w != nullptr &&
                    k_ln_w != nullptr && q_scale != nullptr,
                "timesfm_attention_bwd: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && num_patches > 0 && num_heads > 0, "timesfm_attention_bwd: bad sizes");
  TSFMX_REQUIRE((qkv_dtype == TSFMX_DT_F32 || qkv_dtype == TSFMX_DT_BF16) &&
                    (dout_dtype == TSFMX_DT_F32 || dout_dtype == TSFMX_DT_BF16),
                "timesfm_attention_bwd: qkv / d_out must be f32 or bf16");
  TSFMX_REQUIRE(dqkv_dtype >= TSFMX_DT_F32 && dqkv_dtype <= TSFMX_DT_BF16_SPLIT, "timesfm_attention_bwd: bad dqkv_dtype");
  if (head_dim != 80) {
    set_error("timesfm_attention_bwd: head_dim %d unsupported (TimesFM 2.5 uses 80)", head_dim);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  const int N = num_patches;
  const bool aligned = reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(d_out) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(dqkv) % 16 == 0;
  if (qkv_dtype == TSFMX_DT_BF16 && dout_dtype == TSFMX_DT_BF16 && dqkv_dtype == TSFMX_DT_BF16 && aligned && N <= 64 &&
      !g_force_simt_attention)
    return launch_timesfm_attention_bwd_mma(qkv, d_out, batch, N, num_heads, patch_mask, num_masked, inv_freq, q_ln_w,
                                            k_ln_w, q_scale, eps, dqkv, dparams, stream);
  const int per_warp_bytes = (9 * N * 81 + 4 * N) * 4;
  int wpb = (200 * 1024) / per_warp_bytes;
  if (wpb > 4) wpb = 4;
  if (wpb < 1) {
    set_error("timesfm_attention_bwd: %d patches need %d bytes of shared memory per warp; unsupported", N,
              per_warp_bytes);
    return TSFMX_ERR_UNSUPPORTED;
  }
  const int smem = wpb * per_warp_bytes;
  const int64_t total = batch * num_heads;
  const int64_t blocks = (total + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("timesfm_attention_bwd: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
    kern<<<grid, wpb * 32, smem, stream>>>(qkv, qkv_dtype, d_out, dout_dtype, batch, N, num_heads, patch_mask, num_masked,
                                           inv_freq, q_ln_w, k_ln_w, q_scale, eps, dqkv, dparams);
    return check_last_launch("timesfm_attention_bwd");
  };
  if (dqkv_dtype == TSFMX_DT_F32) return launch(timesfm_attention_bwd_kernel<80, TSFMX_DT_F32>);
  if (dqkv_dtype == TSFMX_DT_BF16) return launch(timesfm_attention_bwd_kernel<80, TSFMX_DT_BF16>);
  return launch(timesfm_attention_bwd_kernel<80, TSFMX_DT_BF16_SPLIT>);
}

extern "C" int tsfmx_transpose_mask(const void* in, int32_t in_dtype, int64_t rows, int32_t cols, int64_t ld_in,
                                    const void* mask, int32_t mask_dtype, int64_t ld_mask, int32_t out_dtype, void* out,
                                    int64_t ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(in != nullptr && out != nullptr, "transpose_mask: NULL pointer");
  TSFMX_REQUIRE(rows > 0 && cols > 0 && ld_out >= rows, "transpose_mask: bad sizes");
  TSFMX_REQUIRE(in_dtype >= TSFMX_DT_F32 && in_dtype <= TSFMX_DT_BF16_SPLIT, "transpose_mask: bad in_dtype");
  TSFMX_REQUIRE(out_dtype == TSFMX_DT_BF16 || out_dtype == TSFMX_DT_BF16_SPLIT, "transpose_mask: out must be bf16 / split");
  TSFMX_REQUIRE(mask == nullptr || (mask_dtype >= TSFMX_DT_F32 && mask_dtype <= TSFMX_DT_BF16_SPLIT),
                "transpose_mask: bad mask_dtype");
  auto pairs_ok = [](const void* p, int dtype, int64_t ld) {
    const int e = dtype == TSFMX_DT_F32 ? 4 : 2;
    return reinterpret_cast<uintptr_t>(p) % (2 * e) == 0 && ld % 2 == 0;
  };
  if (cols % 2 == 0 && ld_out % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0 && pairs_ok(in, in_dtype, ld_in) &&
      (mask == nullptr || pairs_ok(mask, mask_dtype, ld_mask))) {
    const dim3 grid64(static_cast<unsigned>((ld_out + 63) / 64), static_cast<unsigned>((cols + 63) / 64));
    if (out_dtype == TSFMX_DT_BF16)
      transpose_mask64_kernel<TSFMX_DT_BF16><<<grid64, 256, 0, stream>>>(in, in_dtype, rows, cols, ld_in, mask, mask_dtype,
                                                                        ld_mask, out, ld_out);
    else
      transpose_mask64_kernel<TSFMX_DT_BF16_SPLIT><<<grid64, 256, 0, stream>>>(in, in_dtype, rows, cols, ld_in, mask,
                                                                              mask_dtype, ld_mask, out, ld_out);
    return check_last_launch("transpose_mask");
  }
  const dim3 grid(static_cast<unsigned>((ld_out + 31) / 32), static_cast<unsigned>((cols + 31) / 32));
  const dim3 block(32, 8);
  if (out_dtype == TSFMX_DT_BF16)
    transpose_mask_kernel<TSFMX_DT_BF16><<<grid, block, 0, stream>>>(in, in_dtype, rows, cols, ld_in, mask, mask_dtype,
                                                                    ld_mask, out, ld_out);
  else
    transpose_mask_kernel<TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(in, in_dtype, rows, cols, ld_in, mask,
                                                                          mask_dtype, ld_mask, out, ld_out);
  return check_last_launch("transpose_mask");
}

extern "C" int tsfmx_mask_cast_rows(const float* in, int64_t rows, int32_t cols, const void* mask, int32_t mask_dtype,
                                    int64_t ld_mask, int32_t out_dtype, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(in != nullptr && mask != nullptr && out != nullptr, "mask_cast_rows: NULL pointer");
  TSFMX_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "mask_cast_rows: cols must be a multiple of 4");
  TSFMX_REQUIRE(mask_dtype >= TSFMX_DT_F32 && mask_dtype <= TSFMX_DT_BF16_SPLIT, "mask_cast_rows: bad mask_dtype");
  if (rows == 0) return TSFMX_OK;
  const int64_t total = rows * (cols / 4);
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
  if (out_dtype == TSFMX_DT_BF16)
    mask_cast_rows_kernel<TSFMX_DT_BF16><<<grid, 256, 0, stream>>>(in, rows, cols, mask, mask_dtype, ld_mask, out);
  else if (out_dtype == TSFMX_DT_BF16_SPLIT)
    mask_cast_rows_kernel<TSFMX_DT_BF16_SPLIT><<<grid, 256, 0, stream>>>(in, rows, cols, mask, mask_dtype, ld_mask, out);
  else if (out_dtype == TSFMX_DT_F32)
    mask_cast_rows_kernel<TSFMX_DT_F32><<<grid, 256, 0, stream>>>(in, rows, cols, mask, mask_dtype, ld_mask, out);
  else {
    set_error("mask_cast_rows: bad out_dtype %d", out_dtype);
    return TSFMX_ERR_INVALID_ARGUMENT;
  }
  return check_last_launch("mask_cast_rows");
}

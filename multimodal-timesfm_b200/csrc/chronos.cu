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

// ----------------------------------------------------------------------------------------
// Tensor-core version for the throughput mode (bf16 qkv in, bf16 out), T <= 16 * NT tokens.
//
// One CTA of 7 warps per (series, head).  The head's q / k rows are fetched with 16-byte loads, rotated in fp32 with
// a precomputed (cos, sin) table and parked as bf16 in padded shared-memory tiles; v rows go there with cp.async.
// Every warp then owns 16-query tiles: S = Q K^T on mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with the whole score row
// block (T <= 208 keys) in registers - no online-softmax rescaling - key mask and softmax in registers, O = P V on
// the tensor cores again, and the result leaves through the tile's own (dead) Q rows as 16-byte coalesced stores.
// tcgen05 is deliberately not used: per (series, head) the problem is 97..193 x 64, far below its 128-row tiles.
// ----------------------------------------------------------------------------------------
constexpr int ENC_HD = 64;
constexpr int ENC_LD = 72;  // padded row (144 B): ldmatrix rows land in distinct bank groups
constexpr int ENC_WARPS = 7;

__device__ __forceinline__ void enc_cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void enc_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void enc_ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void enc_mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, int half, int seq, float2* __restrict__ table) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < seq * half; i += gridDim.x * blockDim.x) {
    const int t = i / half, d = i - t * half;
    float sn, cs;
    sincosf(static_cast<float>(t) * __ldg(inv_freq + d), &sn, &cs);
    table[i] = make_float2(cs, sn);
  }
}

template <int NT>
__global__ void __launch_bounds__(ENC_WARPS * 32, (NT <= 9 ? 2 : 1)) encoder_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, int seq, int num_heads, const uint8_t* __restrict__ key_mask,
    const float2* __restrict__ rope, __nv_bfloat16* __restrict__ out) {
  constexpr int ROWS = 16 * NT;
  constexpr int HALF = ENC_HD / 2;
  constexpr int TILE = ROWS * ENC_LD;
  extern __shared__ __align__(16) uint8_t smem_enc[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_enc);
  __nv_bfloat16* sK = sQ + TILE;
  __nv_bfloat16* sV = sK + TILE;
  float* s_bias = reinterpret_cast<float*>(sV + TILE);  // [ROWS] additive key mask: 0 = may be attended, -inf = not
  __shared__ int s_any;
  const int T = seq;
  const int b = blockIdx.x / num_heads, h = blockIdx.x - b * num_heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = num_heads * ENC_HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const __nv_bfloat16* gbase = qkv + static_cast<int64_t>(b) * T * ld + h * ENC_HD;

  if (threadIdx.x == 0) s_any = 0;
  // ---- v rows: raw 16-byte async copies (8 per row); rows >= T zero-filled for all three tiles
  for (int c = threadIdx.x; c < T * 8; c += blockDim.x) {
    const int row = c >> 3, ch = c & 7;
    enc_cp_async_16(sV + row * ENC_LD + ch * 8, gbase + row * ld + 2 * width + ch * 8);
  }
  for (int c = threadIdx.x; c < (ROWS - T) * 8 * 3; c += blockDim.x) {
    const int which = c / ((ROWS - T) * 8), rem = c - which * ((ROWS - T) * 8);
    const int row = T + (rem >> 3), ch = rem & 7;
    *reinterpret_cast<uint4*>(sQ + which * TILE + row * ENC_LD + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();  // s_any = 0 visible before the ORs below
  {
    int any = 0;
    for (int j = threadIdx.x; j < ROWS; j += blockDim.x) {
      const int v = j < T && (key_mask == nullptr || key_mask[static_cast<int64_t>(b) * T + j] != 0) ? 1 : 0;
      s_bias[j] = v ? 0.f : -INFINITY;
      any |= v;
    }
    if (any) s_any = 1;  // benign race: every writer stores 1
  }
  // ---- q / k rows with the rotary embedding (rotate-half, position = token index): one thread per (row, 8 dims of
  //      the first half + the matching 8 dims of the second half)
  for (int c = threadIdx.x; c < T * 4; c += blockDim.x) {
    const int row = c >> 2, ch = c & 3;
    const __nv_bfloat16* g = gbase + row * ld + ch * 8;
    const uint4 qa = *reinterpret_cast<const uint4*>(g), qb = *reinterpret_cast<const uint4*>(g + HALF);
    const uint4 ka = *reinterpret_cast<const uint4*>(g + width), kb = *reinterpret_cast<const uint4*>(g + width + HALF);
    const float4* rp = reinterpret_cast<const float4*>(rope + row * HALF + ch * 8);
    const uint32_t qaw[4] = {qa.x, qa.y, qa.z, qa.w}, qbw[4] = {qb.x, qb.y, qb.z, qb.w};
    const uint32_t kaw[4] = {ka.x, ka.y, ka.z, ka.w}, kbw[4] = {kb.x, kb.y, kb.z, kb.w};
    uint32_t oq1[4], oq2[4], ok1[4], ok2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 cs = __ldg(rp + i);  // (cos, sin) of dims 2i, 2i + 1
      const float q1a = bf16_lo(qaw[i]), q1b = bf16_hi(qaw[i]), q2a = bf16_lo(qbw[i]), q2b = bf16_hi(qbw[i]);
      const float k1a = bf16_lo(kaw[i]), k1b = bf16_hi(kaw[i]), k2a = bf16_lo(kbw[i]), k2b = bf16_hi(kbw[i]);
      oq1[i] = pack_bf16x2(q1a * cs.x - q2a * cs.y, q1b * cs.z - q2b * cs.w);
      oq2[i] = pack_bf16x2(q2a * cs.x + q1a * cs.y, q2b * cs.z + q1b * cs.w);
      ok1[i] = pack_bf16x2(k1a * cs.x - k2a * cs.y, k1b * cs.z - k2b * cs.w);
      ok2[i] = pack_bf16x2(k2a * cs.x + k1a * cs.y, k2b * cs.z + k1b * cs.w);
    }
    *reinterpret_cast<uint4*>(sQ + row * ENC_LD + ch * 8) = make_uint4(oq1[0], oq1[1], oq1[2], oq1[3]);
    *reinterpret_cast<uint4*>(sQ + row * ENC_LD + HALF + ch * 8) = make_uint4(oq2[0], oq2[1], oq2[2], oq2[3]);
    *reinterpret_cast<uint4*>(sK + row * ENC_LD + ch * 8) = make_uint4(ok1[0], ok1[1], ok1[2], ok1[3]);
    *reinterpret_cast<uint4*>(sK + row * ENC_LD + HALF + ch * 8) = make_uint4(ok2[0], ok2[1], ok2[2], ok2[3]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const bool has_key = s_any != 0;  // every key masked: uniform weights over all T keys (finfo.min semantics)
  if (!has_key) {
    for (int j = threadIdx.x; j < ROWS; j += blockDim.x) s_bias[j] = j < T ? 0.f : -INFINITY;
    __syncthreads();
  }
  const int g = lane >> 2, t = lane & 3;
  const int ntk = (T + 15) >> 4;
  const uint32_t sq_addr = smem_u32(sQ), sk_addr = smem_u32(sK), sv_addr = smem_u32(sV);
  for (int qi = warp; qi < ntk; qi += ENC_WARPS) {
    uint32_t qf[4][4];
    {
      const int row = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int col = 8 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) enc_ldmatrix_x4(sq_addr + (row * ENC_LD + col + 16 * ks) * 2, qf[ks]);
    }
    float s[NT][2][4];
#pragma unroll
    for (int kj = 0; kj < NT; ++kj) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[kj][nt][e] = 0.f;
      if (kj < ntk) {
        const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
        const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kb[4];
          enc_ldmatrix_x4(sk_addr + (key * ENC_LD + col + 16 * ks) * 2, kb);
          enc_mma_16816(s[kj][0], qf[ks], kb[0], kb[1]);
          enc_mma_16816(s[kj][1], qf[ks], kb[2], kb[3]);
        }
      }
    }
    // mask + row max (rows g and g + 8 of the tile; this thread holds keys 2t, 2t + 1 of every 8-key group): the key
    // mask is an additive 0 / -inf row in shared memory (one 8-byte load and four adds per 8-key group; the select
    // chains it replaces were 19 % of the kernel's instructions).  Every key masked (rare, block-uniform): the
    // staging phase rewrote the mask row to "every existing key" and the scores count as 0, i.e. uniform weights.
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int kj = 0; kj < NT; ++kj)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int key = kj * 16 + nt * 8 + 2 * t;
        const float2 bias = *reinterpret_cast<const float2*>(s_bias + key);
        s[kj][nt][0] = (has_key ? s[kj][nt][0] : 0.f) + bias.x, s[kj][nt][1] = (has_key ? s[kj][nt][1] : 0.f) + bias.y;
        s[kj][nt][2] = (has_key ? s[kj][nt][2] : 0.f) + bias.x, s[kj][nt][3] = (has_key ? s[kj][nt][3] : 0.f) + bias.y;
        mx0 = fmaxf(mx0, fmaxf(s[kj][nt][0], s[kj][nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[kj][nt][2], s[kj][nt][3]));
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    const float ml0 = mx0 * LOG2E_F, ml1 = mx1 * LOG2E_F;  // finite: at least one key of the row carries a score
#pragma unroll
    for (int kj = 0; kj < NT; ++kj)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = exp_sub(s[kj][nt][e], (e & 2) ? ml1 : ml0);
          s[kj][nt][e] = p;
          if (e & 2) sum1 += p; else sum0 += p;
        }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
    for (int kj = 0; kj < NT; ++kj) {
      if (kj < ntk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[kj][0][0], s[kj][0][1]);
        pa[1] = pack_bf16x2(s[kj][0][2], s[kj][0][3]);
        pa[2] = pack_bf16x2(s[kj][1][0], s[kj][1][1]);
        pa[3] = pack_bf16x2(s[kj][1][2], s[kj][1][3]);
        const int key = kj * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int col = 8 * (lane >> 4);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vb[4];
          enc_ldmatrix_x4_trans(sv_addr + (key * ENC_LD + col + 16 * dp) * 2, vb);
          enc_mma_16816(o[2 * dp], pa, vb[0], vb[1]);
          enc_mma_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    // the Q rows of this tile are dead now (only this warp read them): stage the output there
    __syncwarp();
    const int row0 = qi * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      *reinterpret_cast<uint32_t*>(sQ + row0 * ENC_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
      *reinterpret_cast<uint32_t*>(sQ + row1 * ENC_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
    __syncwarp();
    __nv_bfloat16* obase = out + static_cast<int64_t>(b) * T * width + h * ENC_HD;
    for (int c = lane; c < 16 * 8; c += 32) {
      const int row = qi * 16 + (c >> 3), ch = c & 7;
      if (row < T)
        *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * width + ch * 8) =
            *reinterpret_cast<const uint4*>(sQ + row * ENC_LD + ch * 8);
    }
  }
}

template <int NT>
int launch_encoder_attention_mma(const void* qkv, int64_t batch, int seq, int num_heads, const uint8_t* key_mask,
                                 const float* rope, void* out, cudaStream_t stream) {
  constexpr int smem = 3 * 16 * NT * ENC_LD * 2 + 16 * NT * 4;
  auto kern = encoder_attention_mma_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("encoder_attention_mma: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
  }
  kern<<<static_cast<int>(batch * num_heads), ENC_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), seq, num_heads, key_mask, reinterpret_cast<const float2*>(rope),
      reinterpret_cast<__nv_bfloat16*>(out));
  return check_last_launch("encoder_attention_mma");
}

// ----------------------------------------------------------------------------------------
// Tensor-core backward of the encoder attention core (bf16 qkv / dO in, bf16 dqkv out, T <= 112): one CTA of 7 warps
// per (series, head), same staging as the forward kernel plus dO.
//   pass A  warp w owns query tile w: S = QK^T and dP = dO V^T on mma.sync, softmax / delta / dS in registers,
//           dQ = dS K on the tensor cores, RoPE transposed in registers (a thread's C fragment holds dims d and d + 32
//           of its rows), row statistics to shared memory
//   pass B  warp w owns key tile w: P and dS blocks recomputed per query tile, transposed in registers (movmatrix),
//           dV = P^T dO and dK = dS^T Q; dK un-rotated in registers
// Results are staged as bf16 tiles and leave with 16-byte coalesced stores.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t enc_movmatrix(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int NT>
__global__ void __launch_bounds__(ENC_WARPS * 32, 1) encoder_attention_bwd_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, int seq, int num_heads,
    const uint8_t* __restrict__ key_mask, const float2* __restrict__ rope, __nv_bfloat16* __restrict__ dqkv) {
  constexpr int ROWS = 16 * NT;
  constexpr int HALF = ENC_HD / 2;
  constexpr int TILE = ROWS * ENC_LD;
  extern __shared__ __align__(16) uint8_t smem_encb[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_encb);
  __nv_bfloat16* sK = sQ + TILE;
  __nv_bfloat16* sV = sK + TILE;
  __nv_bfloat16* sDO = sV + TILE;
  __nv_bfloat16* sDQ = sDO + TILE;
  __nv_bfloat16* sDK = sDQ + TILE;
  __nv_bfloat16* sDV = sDK + TILE;
  float* sMx = reinterpret_cast<float*>(sDV + TILE);
  float* sInv = sMx + ROWS;
  float* sDelta = sInv + ROWS;
  uint8_t* s_valid = reinterpret_cast<uint8_t*>(sDelta + ROWS);
  __shared__ int s_any;
  const int T = seq;
  const int b = blockIdx.x / num_heads, h = blockIdx.x - b * num_heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = num_heads * ENC_HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const __nv_bfloat16* gbase = qkv + static_cast<int64_t>(b) * T * ld + h * ENC_HD;
  const __nv_bfloat16* dobase = dout + static_cast<int64_t>(b) * T * width + h * ENC_HD;

  if (threadIdx.x == 0) s_any = 0;
  for (int c = threadIdx.x; c < T * 8; c += blockDim.x) {
    const int row = c >> 3, ch = c & 7;
    enc_cp_async_16(sV + row * ENC_LD + ch * 8, gbase + row * ld + 2 * width + ch * 8);
    enc_cp_async_16(sDO + row * ENC_LD + ch * 8, dobase + static_cast<int64_t>(row) * width + ch * 8);
  }
  for (int c = threadIdx.x; c < (ROWS - T) * 8 * 4; c += blockDim.x) {
    const int which = c / ((ROWS - T) * 8), rem = c - which * ((ROWS - T) * 8);
    const int row = T + (rem >> 3), ch = rem & 7;
    *reinterpret_cast<uint4*>(sQ + which * TILE + row * ENC_LD + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  {
    int any = 0;
    for (int j = threadIdx.x; j < ROWS; j += blockDim.x) {
      const uint8_t v = j < T && (key_mask == nullptr || key_mask[static_cast<int64_t>(b) * T + j] != 0) ? 1 : 0;
      s_valid[j] = v;
      any |= v;
    }
    if (any) s_any = 1;
  }
  for (int c = threadIdx.x; c < T * 4; c += blockDim.x) {
    const int row = c >> 2, ch = c & 3;
    const __nv_bfloat16* g = gbase + row * ld + ch * 8;
    const uint4 qa = *reinterpret_cast<const uint4*>(g), qb = *reinterpret_cast<const uint4*>(g + HALF);
    const uint4 ka = *reinterpret_cast<const uint4*>(g + width), kb = *reinterpret_cast<const uint4*>(g + width + HALF);
    const float4* rp = reinterpret_cast<const float4*>(rope + row * HALF + ch * 8);
    const uint32_t qaw[4] = {qa.x, qa.y, qa.z, qa.w}, qbw[4] = {qb.x, qb.y, qb.z, qb.w};
    const uint32_t kaw[4] = {ka.x, ka.y, ka.z, ka.w}, kbw[4] = {kb.x, kb.y, kb.z, kb.w};
    uint32_t oq1[4], oq2[4], ok1[4], ok2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 cs = __ldg(rp + i);
      const float q1a = bf16_lo(qaw[i]), q1b = bf16_hi(qaw[i]), q2a = bf16_lo(qbw[i]), q2b = bf16_hi(qbw[i]);
      const float k1a = bf16_lo(kaw[i]), k1b = bf16_hi(kaw[i]), k2a = bf16_lo(kbw[i]), k2b = bf16_hi(kbw[i]);
      oq1[i] = pack_bf16x2(q1a * cs.x - q2a * cs.y, q1b * cs.z - q2b * cs.w);
      oq2[i] = pack_bf16x2(q2a * cs.x + q1a * cs.y, q2b * cs.z + q1b * cs.w);
      ok1[i] = pack_bf16x2(k1a * cs.x - k2a * cs.y, k1b * cs.z - k2b * cs.w);
      ok2[i] = pack_bf16x2(k2a * cs.x + k1a * cs.y, k2b * cs.z + k1b * cs.w);
    }
    *reinterpret_cast<uint4*>(sQ + row * ENC_LD + ch * 8) = make_uint4(oq1[0], oq1[1], oq1[2], oq1[3]);
    *reinterpret_cast<uint4*>(sQ + row * ENC_LD + HALF + ch * 8) = make_uint4(oq2[0], oq2[1], oq2[2], oq2[3]);
    *reinterpret_cast<uint4*>(sK + row * ENC_LD + ch * 8) = make_uint4(ok1[0], ok1[1], ok1[2], ok1[3]);
    *reinterpret_cast<uint4*>(sK + row * ENC_LD + HALF + ch * 8) = make_uint4(ok2[0], ok2[1], ok2[2], ok2[3]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const bool has_key = s_any != 0;
  const int g = lane >> 2, t = lane & 3;
  const int ntk = (T + 15) >> 4;
  const uint32_t sq_addr = smem_u32(sQ), sk_addr = smem_u32(sK), sv_addr = smem_u32(sV), sdo_addr = smem_u32(sDO);
  const float inv_T = 1.0f / static_cast<float>(T);

  // un-rotate a 16 x 64 fp32 C-fragment block (rows row0 = base + g and row0 + 8) and park it as bf16
  auto unrotate_store = [&](float (&acc)[8][4], int tile_row0, __nv_bfloat16* dst) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int row = tile_row0 + g + 8 * hrow;
      const int rr = row < T ? row : 0;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(rope + rr * HALF + dt * 8 + 2 * t));  // dims d, d + 1
        const float a0 = acc[dt][2 * hrow], a1 = acc[dt][2 * hrow + 1];          // dims d, d + 1
        const float b0 = acc[dt + 4][2 * hrow], b1 = acc[dt + 4][2 * hrow + 1];  // dims d + 32, d + 33
        *reinterpret_cast<uint32_t*>(dst + row * ENC_LD + dt * 8 + 2 * t) = pack_bf16x2(a0 * cs.x + b0 * cs.y, a1 * cs.z + b1 * cs.w);
        *reinterpret_cast<uint32_t*>(dst + row * ENC_LD + HALF + dt * 8 + 2 * t) =
            pack_bf16x2(b0 * cs.x - a0 * cs.y, b1 * cs.z - a1 * cs.w);
      }
    }
  };

  // ---- pass A: per query tile
  for (int qi = warp; qi < ntk; qi += ENC_WARPS) {
    uint32_t qf[4][4], df[4][4];
    {
      const int row = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int col = 8 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        enc_ldmatrix_x4(sq_addr + (row * ENC_LD + col + 16 * ks) * 2, qf[ks]);
        enc_ldmatrix_x4(sdo_addr + (row * ENC_LD + col + 16 * ks) * 2, df[ks]);
      }
    }
    float s[NT][2][4], dp[NT][2][4];
#pragma unroll
    for (int kj = 0; kj < NT; ++kj) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[kj][nt][e] = 0.f, dp[kj][nt][e] = 0.f;
      if (kj < ntk) {
        const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
        const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kb[4], vb[4];
          enc_ldmatrix_x4(sk_addr + (key * ENC_LD + col + 16 * ks) * 2, kb);
          enc_mma_16816(s[kj][0], qf[ks], kb[0], kb[1]);
          enc_mma_16816(s[kj][1], qf[ks], kb[2], kb[3]);
          enc_ldmatrix_x4(sv_addr + (key * ENC_LD + col + 16 * ks) * 2, vb);
          enc_mma_16816(dp[kj][0], df[ks], vb[0], vb[1]);
          enc_mma_16816(dp[kj][1], df[ks], vb[2], vb[3]);
        }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int kj = 0; kj < NT; ++kj)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int key = kj * 16 + nt * 8 + 2 * t;
        const uint32_t vv = *reinterpret_cast<const uint16_t*>(s_valid + key);
        const bool ok0 = has_key ? (vv & 0xffu) != 0 : key < T;
        const bool ok1 = has_key ? (vv >> 8) != 0 : key + 1 < T;
        s[kj][nt][0] = ok0 ? (has_key ? s[kj][nt][0] : 0.f) : -INFINITY;
        s[kj][nt][1] = ok1 ? (has_key ? s[kj][nt][1] : 0.f) : -INFINITY;
        s[kj][nt][2] = ok0 ? (has_key ? s[kj][nt][2] : 0.f) : -INFINITY;
        s[kj][nt][3] = ok1 ? (has_key ? s[kj][nt][3] : 0.f) : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[kj][nt][0], s[kj][nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[kj][nt][2], s[kj][nt][3]));
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    const float ml0 = mx0 * LOG2E_F, ml1 = mx1 * LOG2E_F;  // row maxima in the exp2 domain (kept for pass B)
#pragma unroll
    for (int kj = 0; kj < NT; ++kj)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = exp_sub(s[kj][nt][e], (e & 2) ? ml1 : ml0);  // masked scores are -inf -> 0
          s[kj][nt][e] = p;
          if (e & 2) sum1 += p; else sum0 += p;
        }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int kj = 0; kj < NT; ++kj)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = s[kj][nt][e] * ((e & 2) ? inv1 : inv0);
          s[kj][nt][e] = p;
          if (e & 2) dl1 += p * dp[kj][nt][e]; else dl0 += p * dp[kj][nt][e];
        }
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1);
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
    const int row0 = qi * 16 + g, row1 = row0 + 8;
    if (t == 0) {
      sMx[row0] = ml0, sInv[row0] = inv0, sDelta[row0] = dl0;
      sMx[row1] = ml1, sInv[row1] = inv1, sDelta[row1] = dl1;
    }
    float dq[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[dt][e] = 0.f;
#pragma unroll
    for (int kj = 0; kj < NT; ++kj) {
      if (kj < ntk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[kj][0][0] * (dp[kj][0][0] - dl0), s[kj][0][1] * (dp[kj][0][1] - dl0));
        pa[1] = pack_bf16x2(s[kj][0][2] * (dp[kj][0][2] - dl1), s[kj][0][3] * (dp[kj][0][3] - dl1));
        pa[2] = pack_bf16x2(s[kj][1][0] * (dp[kj][1][0] - dl0), s[kj][1][1] * (dp[kj][1][1] - dl0));
        pa[3] = pack_bf16x2(s[kj][1][2] * (dp[kj][1][2] - dl1), s[kj][1][3] * (dp[kj][1][3] - dl1));
        const int key = kj * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int col = 8 * (lane >> 4);
#pragma unroll
        for (int dpp = 0; dpp < 4; ++dpp) {
          uint32_t kb[4];
          enc_ldmatrix_x4_trans(sk_addr + (key * ENC_LD + col + 16 * dpp) * 2, kb);
          enc_mma_16816(dq[2 * dpp], pa, kb[0], kb[1]);
          enc_mma_16816(dq[2 * dpp + 1], pa, kb[2], kb[3]);
        }
      }
    }
    unrotate_store(dq, qi * 16, sDQ);
  }
  __syncthreads();

  // ---- pass B: per key tile
  for (int kj = warp; kj < ntk; kj += ENC_WARPS) {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dv[dt][e] = 0.f, dk[dt][e] = 0.f;
    uint32_t kf[4][4], vf[4][4];
    {
      const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
      const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        enc_ldmatrix_x4(sk_addr + (key * ENC_LD + col + 16 * ks) * 2, kf[ks]);
        enc_ldmatrix_x4(sv_addr + (key * ENC_LD + col + 16 * ks) * 2, vf[ks]);
      }
    }
    bool okk[2][2];  // [nt][e & 1]: validity of this thread's keys in the tile
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = kj * 16 + nt * 8 + 2 * t;
      const uint32_t vv = *reinterpret_cast<const uint16_t*>(s_valid + key);
      okk[nt][0] = has_key ? (vv & 0xffu) != 0 : key < T;
      okk[nt][1] = has_key ? (vv >> 8) != 0 : key + 1 < T;
    }
#pragma unroll 1
    for (int qi = 0; qi < ntk; ++qi) {
      float s2[2][4], dp2[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s2[nt][e] = 0.f, dp2[nt][e] = 0.f;
      const int qrow = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int qcol = 8 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qa[4], da[4];
        enc_ldmatrix_x4(sq_addr + (qrow * ENC_LD + qcol + 16 * ks) * 2, qa);
        enc_ldmatrix_x4(sdo_addr + (qrow * ENC_LD + qcol + 16 * ks) * 2, da);
        enc_mma_16816(s2[0], qa, kf[ks][0], kf[ks][1]);
        enc_mma_16816(s2[1], qa, kf[ks][2], kf[ks][3]);
        enc_mma_16816(dp2[0], da, vf[ks][0], vf[ks][1]);
        enc_mma_16816(dp2[1], da, vf[ks][2], vf[ks][3]);
      }
      const int row0 = qi * 16 + g, row1 = row0 + 8;
      const float mx0 = sMx[row0], mx1 = sMx[row1], inv0 = sInv[row0], inv1 = sInv[row1];
      const float dl0 = sDelta[row0], dl1 = sDelta[row1];
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e & 2) ? row1 : row0;
          float pv = 0.f;
          if (okk[nt][e & 1] && row < T)
            pv = has_key ? exp_sub(s2[nt][e], (e & 2) ? mx1 : mx0) * ((e & 2) ? inv1 : inv0) : inv_T;  // sMx: mx * log2e
          p[nt][e] = pv;
          ds[nt][e] = pv * (dp2[nt][e] - ((e & 2) ? dl1 : dl0));
        }
      uint32_t pT[4], dsT[4];
      pT[0] = enc_movmatrix(pack_bf16x2(p[0][0], p[0][1]));
      pT[1] = enc_movmatrix(pack_bf16x2(p[1][0], p[1][1]));
      pT[2] = enc_movmatrix(pack_bf16x2(p[0][2], p[0][3]));
      pT[3] = enc_movmatrix(pack_bf16x2(p[1][2], p[1][3]));
      dsT[0] = enc_movmatrix(pack_bf16x2(ds[0][0], ds[0][1]));
      dsT[1] = enc_movmatrix(pack_bf16x2(ds[1][0], ds[1][1]));
      dsT[2] = enc_movmatrix(pack_bf16x2(ds[0][2], ds[0][3]));
      dsT[3] = enc_movmatrix(pack_bf16x2(ds[1][2], ds[1][3]));
#pragma unroll
      for (int dpp = 0; dpp < 4; ++dpp) {
        uint32_t ob[4], qb[4];
        enc_ldmatrix_x4_trans(sdo_addr + (qrow * ENC_LD + qcol + 16 * dpp) * 2, ob);
        enc_mma_16816(dv[2 * dpp], pT, ob[0], ob[1]);
        enc_mma_16816(dv[2 * dpp + 1], pT, ob[2], ob[3]);
        enc_ldmatrix_x4_trans(sq_addr + (qrow * ENC_LD + qcol + 16 * dpp) * 2, qb);
        enc_mma_16816(dk[2 * dpp], dsT, qb[0], qb[1]);
        enc_mma_16816(dk[2 * dpp + 1], dsT, qb[2], qb[3]);
      }
    }
    unrotate_store(dk, kj * 16, sDK);
    const int key0 = kj * 16 + g, key1 = key0 + 8;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      *reinterpret_cast<uint32_t*>(sDV + key0 * ENC_LD + dt * 8 + 2 * t) = pack_bf16x2(dv[dt][0], dv[dt][1]);
      *reinterpret_cast<uint32_t*>(sDV + key1 * ENC_LD + dt * 8 + 2 * t) = pack_bf16x2(dv[dt][2], dv[dt][3]);
    }
  }
  __syncthreads();
  __nv_bfloat16* obase = dqkv + static_cast<int64_t>(b) * T * ld + h * ENC_HD;
  for (int c = threadIdx.x; c < T * 8; c += blockDim.x) {
    const int row = c >> 3, ch = c & 7;
    __nv_bfloat16* o = obase + row * ld + ch * 8;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(sDQ + row * ENC_LD + ch * 8);
    *reinterpret_cast<uint4*>(o + width) = *reinterpret_cast<const uint4*>(sDK + row * ENC_LD + ch * 8);
    *reinterpret_cast<uint4*>(o + 2 * width) = *reinterpret_cast<const uint4*>(sDV + row * ENC_LD + ch * 8);
  }
}

template <int NT>
int launch_encoder_attention_bwd_mma(const void* qkv, const void* dout, int64_t batch, int seq, int num_heads,
                                     const uint8_t* key_mask, const float* rope, void* dqkv, cudaStream_t stream) {
  constexpr int smem = 7 * 16 * NT * ENC_LD * 2 + 3 * 16 * NT * 4 + 16 * NT;
  auto kern = encoder_attention_bwd_mma_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("encoder_attention_bwd_mma: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
  }
  kern<<<static_cast<int>(batch * num_heads), ENC_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout), seq, num_heads, key_mask,
      reinterpret_cast<const float2*>(rope), reinterpret_cast<__nv_bfloat16*>(dqkv));
  return check_last_launch("encoder_attention_bwd_mma");
}

// ----------------------------------------------------------------------------------------
// Backward of the encoder attention core (fusion fine-tune through the frozen Chronos-2 encoder; the reference trains
// the fusion module with either adapter, scripts/tune_time_mmd_sweep.py:124-126).  fp32 SIMT, any T, exact up to
// summation order: pass A gives one warp to a query row (P, dP, delta, dS -> dq, and the row's softmax statistics
// to a workspace), pass B one warp to a key row (dk, dv from the recomputed P and dS).  Lane l holds head dims l and
// l + 32, i.e. one rotary pair, so RoPE and its transpose are lane-local.
// ----------------------------------------------------------------------------------------
template <int OUT>
__device__ __forceinline__ void enc_store2(void* out, int64_t row, int64_t row_elems, int c, float v0, float v1) {
  // row_elems = 3 * width (columns of one dqkv row); c and c + 32 are the two columns written
  if constexpr (OUT == TSFMX_DT_F32) {
    float* o = reinterpret_cast<float*>(out) + row * row_elems;
    o[c] = v0, o[c + 32] = v1;
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * row_elems;
    o[c] = __float2bfloat16_rn(v0), o[c + 32] = __float2bfloat16_rn(v1);
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * 2 * row_elems;  // [hi(3W) | lo(3W)]
    split_bf16(v0, o[c], o[row_elems + c]);
    split_bf16(v1, o[c + 32], o[row_elems + c + 32]);
  }
}

template <int OUT>
__global__ void __launch_bounds__(128) encoder_attention_bwd_q_kernel(
    const void* __restrict__ qkv, int qkv_dtype, const void* __restrict__ dout, int dout_dtype, int64_t batch, int seq,
    int num_heads, const uint8_t* __restrict__ key_mask, const float2* __restrict__ rope, float4* __restrict__ stats,
    void* dqkv) {
  extern __shared__ float s_enc_bwd[];  // [4 warps][2][T]: probabilities, dP
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = seq;
  float* sp = s_enc_bwd + warp * 2 * T;
  float* sdp = sp + T;
  const int width = num_heads * 64;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const int64_t units = batch * num_heads * T;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * 4 + warp; u < units; u += static_cast<int64_t>(gridDim.x) * 4) {
    const int i = static_cast<int>(u % T);
    const int64_t bh = u / T;
    const int h = static_cast<int>(bh % num_heads);
    const int64_t b = bh / num_heads;
    const int64_t row_i = b * T + i;
    const int64_t base = b * T * ld + h * 64;
    const uint8_t* km = key_mask != nullptr ? key_mask + b * T : nullptr;
    const float2 ri = __ldg(rope + i * 32 + lane);
    const float qa = ld_any(qkv, qkv_dtype, base + i * ld + lane), qb = ld_any(qkv, qkv_dtype, base + i * ld + lane + 32);
    const float q0 = qa * ri.x - qb * ri.y, q1 = qb * ri.x + qa * ri.y;
    const float g0 = ld_any(dout, dout_dtype, row_i * width + h * 64 + lane);
    const float g1 = ld_any(dout, dout_dtype, row_i * width + h * 64 + lane + 32);
    float mx = -INFINITY;
    bool any = false;
    for (int j = 0; j < T; ++j) {
      const float2 rj = __ldg(rope + j * 32 + lane);
      const int64_t kr = base + j * ld + width;
      const float ka = ld_any(qkv, qkv_dtype, kr + lane), kb = ld_any(qkv, qkv_dtype, kr + lane + 32);
      const float k0 = ka * rj.x - kb * rj.y, k1 = kb * rj.x + ka * rj.y;
      const float s = warp_sum(q0 * k0 + q1 * k1);
      const float dp = warp_sum(g0 * ld_any(qkv, qkv_dtype, kr + width + lane) + g1 * ld_any(qkv, qkv_dtype, kr + width + lane + 32));
      const bool ok = km == nullptr || km[j] != 0;
      if (lane == 0) sp[j] = ok ? s : -INFINITY, sdp[j] = dp;
      if (ok) mx = fmaxf(mx, s), any = true;
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = any ? (sp[j] == -INFINITY ? 0.f : expf(sp[j] - mx)) : 1.f;  // all keys masked: uniform
      sp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float delta = 0.f;
    for (int j = lane; j < T; j += 32) delta += sp[j] * inv * sdp[j];
    delta = warp_sum(delta);
    float dq0 = 0.f, dq1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = sp[j] * inv * (sdp[j] - delta);
      if (ds != 0.f) {
        const float2 rj = __ldg(rope + j * 32 + lane);
        const int64_t kr = base + j * ld + width;
        const float ka = ld_any(qkv, qkv_dtype, kr + lane), kb = ld_any(qkv, qkv_dtype, kr + lane + 32);
        dq0 = fmaf(ds, ka * rj.x - kb * rj.y, dq0);
        dq1 = fmaf(ds, kb * rj.x + ka * rj.y, dq1);
      }
    }
    // transpose of the rotation
    enc_store2<OUT>(dqkv, row_i, ld, h * 64 + lane, dq0 * ri.x + dq1 * ri.y, dq1 * ri.x - dq0 * ri.y);
    if (lane == 0) stats[(b * num_heads + h) * T + i] = make_float4(mx, inv, delta, any ? 1.f : 0.f);
    __syncwarp();
  }
}

template <int OUT>
__global__ void __launch_bounds__(128) encoder_attention_bwd_kv_kernel(
    const void* __restrict__ qkv, int qkv_dtype, const void* __restrict__ dout, int dout_dtype, int64_t batch, int seq,
    int num_heads, const uint8_t* __restrict__ key_mask, const float2* __restrict__ rope,
    const float4* __restrict__ stats, void* dqkv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = seq;
  const int width = num_heads * 64;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const int64_t units = batch * num_heads * T;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * 4 + warp; u < units; u += static_cast<int64_t>(gridDim.x) * 4) {
    const int j = static_cast<int>(u % T);
    const int64_t bh = u / T;
    const int h = static_cast<int>(bh % num_heads);
    const int64_t b = bh / num_heads;
    const int64_t base = b * T * ld + h * 64;
    const bool ok = key_mask == nullptr || key_mask[b * T + j] != 0;
    const float2 rj = __ldg(rope + j * 32 + lane);
    const int64_t kr = base + j * ld + width;
    const float ka = ld_any(qkv, qkv_dtype, kr + lane), kb = ld_any(qkv, qkv_dtype, kr + lane + 32);
    const float k0 = ka * rj.x - kb * rj.y, k1 = kb * rj.x + ka * rj.y;
    const float v0 = ld_any(qkv, qkv_dtype, kr + width + lane), v1 = ld_any(qkv, qkv_dtype, kr + width + lane + 32);
    float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
    const float4* st = stats + (b * num_heads + h) * T;
    for (int i = 0; i < T; ++i) {
      const float4 si = __ldg(st + i);  // (max, 1 / sum, delta, has_key)
      if (si.w != 0.f && !ok) continue;  // a masked key takes part only in the all-masked (uniform) case
      const float2 ri = __ldg(rope + i * 32 + lane);
      const float qa = ld_any(qkv, qkv_dtype, base + i * ld + lane), qb = ld_any(qkv, qkv_dtype, base + i * ld + lane + 32);
      const float q0 = qa * ri.x - qb * ri.y, q1 = qb * ri.x + qa * ri.y;
      const int64_t orow = (b * T + i) * width + h * 64;
      const float g0 = ld_any(dout, dout_dtype, orow + lane), g1 = ld_any(dout, dout_dtype, orow + lane + 32);
      const float s = warp_sum(q0 * k0 + q1 * k1);
      const float dp = warp_sum(g0 * v0 + g1 * v1);
      const float p = (si.w != 0.f ? expf(s - si.x) : 1.f) * si.y;
      const float ds = p * (dp - si.z);
      dk0 = fmaf(ds, q0, dk0), dk1 = fmaf(ds, q1, dk1);
      dv0 = fmaf(p, g0, dv0), dv1 = fmaf(p, g1, dv1);
    }
    const int64_t row_j = b * T + j;
    enc_store2<OUT>(dqkv, row_j, ld, width + h * 64 + lane, dk0 * rj.x + dk1 * rj.y, dk1 * rj.x - dk0 * rj.y);
    enc_store2<OUT>(dqkv, row_j, ld, 2 * width + h * 64 + lane, dv0, dv1);
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

extern "C" int tsfmx_rope_table(const float* inv_freq, int32_t half_dim, int32_t seq, float* table, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(inv_freq != nullptr && table != nullptr, "rope_table: NULL pointer");
  TSFMX_REQUIRE(half_dim > 0 && seq > 0, "rope_table: bad sizes");
  const int total = half_dim * seq;
  rope_table_kernel<<<(total + 255) / 256, 256, 0, stream>>>(inv_freq, half_dim, seq, reinterpret_cast<float2*>(table));
  return check_last_launch("rope_table");
}

extern "C" int tsfmx_encoder_attention_mma(const void* qkv, int64_t batch, int32_t seq, int32_t num_heads,
                                           int32_t head_dim, const uint8_t* key_mask, const float* rope_table,
                                           void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && out != nullptr && rope_table != nullptr, "encoder_attention_mma: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && seq > 0 && num_heads > 0, "encoder_attention_mma: bad sizes");
  TSFMX_REQUIRE(batch * num_heads < (int64_t(1) << 31), "encoder_attention_mma: too many (series, head) pairs");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(rope_table) % 16 == 0,
                "encoder_attention_mma: pointers must be 16-byte aligned");
  if (head_dim != ENC_HD || seq > 208) {
    set_error("encoder_attention_mma: head_dim %d / seq %d unsupported (head_dim 64, seq <= 208)", head_dim, seq);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  if (seq <= 64) return launch_encoder_attention_mma<4>(qkv, batch, seq, num_heads, key_mask, rope_table, out, stream);
  if (seq <= 112) return launch_encoder_attention_mma<7>(qkv, batch, seq, num_heads, key_mask, rope_table, out, stream);
  if (seq <= 144) return launch_encoder_attention_mma<9>(qkv, batch, seq, num_heads, key_mask, rope_table, out, stream);
  return launch_encoder_attention_mma<13>(qkv, batch, seq, num_heads, key_mask, rope_table, out, stream);
}

extern "C" int tsfmx_encoder_attention_bwd(const void* qkv, int32_t qkv_dtype, const void* d_out, int32_t dout_dtype,
                                           int64_t batch, int32_t seq, int32_t num_heads, int32_t head_dim,
                                           const uint8_t* key_mask, const float* rope_table, float* stats_workspace,
                                           int32_t dqkv_dtype, void* dqkv, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && d_out != nullptr && dqkv != nullptr && rope_table != nullptr && stats_workspace != nullptr,
                "encoder_attention_bwd: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && seq > 0 && num_heads > 0, "encoder_attention_bwd: bad sizes");
  TSFMX_REQUIRE((qkv_dtype == TSFMX_DT_F32 || qkv_dtype == TSFMX_DT_BF16) &&
                    (dout_dtype == TSFMX_DT_F32 || dout_dtype == TSFMX_DT_BF16),
                "encoder_attention_bwd: qkv / d_out must be f32 or bf16");
  TSFMX_REQUIRE(dqkv_dtype >= TSFMX_DT_F32 && dqkv_dtype <= TSFMX_DT_BF16_SPLIT, "encoder_attention_bwd: bad dqkv_dtype");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(stats_workspace) % 16 == 0, "encoder_attention_bwd: workspace must be 16-byte aligned");
  if (head_dim != 64) {
    set_error("encoder_attention_bwd: head_dim %d unsupported (Chronos-2 uses 64)", head_dim);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  {
    auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
    if (qkv_dtype == TSFMX_DT_BF16 && dout_dtype == TSFMX_DT_BF16 && dqkv_dtype == TSFMX_DT_BF16 && seq <= 112 &&
        al16(qkv) && al16(d_out) && al16(dqkv) && al16(rope_table) && !g_force_simt_attention &&
        batch * num_heads < (int64_t(1) << 31)) {
      if (seq <= 64)
        return launch_encoder_attention_bwd_mma<4>(qkv, d_out, batch, seq, num_heads, key_mask, rope_table, dqkv, stream);
      return launch_encoder_attention_bwd_mma<7>(qkv, d_out, batch, seq, num_heads, key_mask, rope_table, dqkv, stream);
    }
  }
  const int smem = 4 * 2 * seq * static_cast<int>(sizeof(float));
  TSFMX_REQUIRE(smem <= 200 * 1024, "encoder_attention_bwd: %d tokens need %d bytes of shared memory", seq, smem);
  const int64_t units = batch * num_heads * seq;
  const int64_t blocks = (units + 3) / 4;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  const float2* rope = reinterpret_cast<const float2*>(rope_table);
  float4* stats = reinterpret_cast<float4*>(stats_workspace);
  auto launch = [&](auto kq, auto kkv) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("encoder_attention_bwd: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
    kq<<<grid, 128, smem, stream>>>(qkv, qkv_dtype, d_out, dout_dtype, batch, seq, num_heads, key_mask, rope, stats, dqkv);
    int rc = check_last_launch("encoder_attention_bwd_q");
    if (rc != TSFMX_OK) return rc;
    kkv<<<grid, 128, 0, stream>>>(qkv, qkv_dtype, d_out, dout_dtype, batch, seq, num_heads, key_mask, rope, stats, dqkv);
    return check_last_launch("encoder_attention_bwd_kv");
  };
  if (dqkv_dtype == TSFMX_DT_F32)
    return launch(encoder_attention_bwd_q_kernel<TSFMX_DT_F32>, encoder_attention_bwd_kv_kernel<TSFMX_DT_F32>);
  if (dqkv_dtype == TSFMX_DT_BF16)
    return launch(encoder_attention_bwd_q_kernel<TSFMX_DT_BF16>, encoder_attention_bwd_kv_kernel<TSFMX_DT_BF16>);
  return launch(encoder_attention_bwd_q_kernel<TSFMX_DT_BF16_SPLIT>, encoder_attention_bwd_kv_kernel<TSFMX_DT_BF16_SPLIT>);
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

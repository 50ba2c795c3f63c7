// Chronos-T5 backbone stages (BASELINE.json north_star / configs[2]: "Chronos-T5-base ... encoder-decoder forecast";
// not part of the reference, which only wraps Chronos-2 - upstream chronos.ChronosModel drives a
// transformers.T5ForConditionalGeneration; HF twin: transformers/models/t5/modeling_t5.py T5Attention.forward):
//
//   embed_rows               token ids -> rows of the (fp32) embedding table
//   t5_attention             general T5 attention core in fp32 SIMT: no 1/sqrt(d) scaling, additive relative-position
//                            bias looked up by (key position - query position), optional causal limit, key mask;
//                            serves the parity mode, the decoder's single-query self / cross attention over the KV
//                            cache, and teacher-forced decoding
//   t5_encoder_attention_mma throughput mode of the encoder self-attention (bf16, T <= 704): one CTA per
//                            (series, head) keeps K and V in shared memory, every warp streams 16-query tiles through
//                            S = QK^T / online softmax / O = PV on mma.sync.m16n8k16
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

constexpr int T5_HD = 64;

// ------------------------------------------------------------------------------------------------ embedding
__global__ void embed_rows_kernel(const int64_t* __restrict__ ids, int64_t rows, int dims, int vocab,
                                  const float* __restrict__ table, float* __restrict__ out) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = dims >> 2;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * warps + warp; r < rows; r += static_cast<int64_t>(gridDim.x) * warps) {
    int64_t id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4* src = reinterpret_cast<const float4*>(table + id * dims);
    float4* dst = reinterpret_cast<float4*>(out + r * dims);
    for (int i = lane; i < nvec; i += 32) dst[i] = __ldg(src + i);
  }
}

// ------------------------------------------------------------------------------------------------ SIMT attention
__device__ __forceinline__ float t5_ld(const void* base, int dtype, int64_t idx) {
  return dtype == TSFMX_DT_F32 ? reinterpret_cast<const float*>(base)[idx]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

struct T5AttnParams {
  const void* q; const void* k; const void* v;
  int q_dtype, kv_dtype;
  int64_t ldq, ldk, ldv;        // row strides in elements
  int64_t q_batch_stride, kv_batch_stride;  // elements between consecutive series
  int kv_batch_div;             // query batch index b reads keys / values / mask of series b / kv_batch_div (sample paths
                                // of one series share the encoder's cross-attention keys and values)
  int tq, tk, num_heads;
  int q_pos0;                   // position of query row 0 (decode step); key j has position j
  int causal;                   // key allowed iff j <= q_pos0 + i
  const uint8_t* key_mask;      // [B, tk] non-zero = attendable, or NULL
  const float* bias;            // [H, bias_len] or NULL; entry for (key pos - query pos) at index delta + bias_zero
  int bias_len, bias_zero;
  void* out; int out_dtype; int64_t ldo; int64_t o_batch_stride;
};

// 64 consecutive elements of a row as floats (16-byte loads; bf16 rows are 128 B, fp32 rows 256 B)
__device__ __forceinline__ void t5_load_row64(const void* base, int dtype, int64_t off, float (&r)[T5_HD]) {
  if (dtype == TSFMX_DT_F32) {
    const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = __ldg(p4 + i);
      r[4 * i] = v.x, r[4 * i + 1] = v.y, r[4 * i + 2] = v.z, r[4 * i + 3] = v.w;
    }
  } else {
    const uint4* p4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 v = __ldg(p4 + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r[8 * i + 2 * e] = __uint_as_float(w[e] << 16);
        r[8 * i + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      }
    }
  }
}

// One warp per (series, query row, head).  Scores: lane l owns keys l, l + 32, ... and reads each key row whole
// (eight independent 16-byte loads in flight per lane, no per-key reduction); P V: lane l owns dims 2l, 2l + 1 and
// streams the value rows (coalesced), four keys per iteration.
template <int OUT>
__global__ void __launch_bounds__(128) t5_attention_kernel(const T5AttnParams p) {
  extern __shared__ float s_scores[];  // [4 warps][tk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = s_scores + warp * p.tk;
  const int64_t units = static_cast<int64_t>(p.tq) * p.num_heads;
  const int b = blockIdx.y;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * 4 + warp; u < units; u += static_cast<int64_t>(gridDim.x) * 4) {
    const int i = static_cast<int>(u / p.num_heads), h = static_cast<int>(u - static_cast<int64_t>(i) * p.num_heads);
    const int qpos = p.q_pos0 + i;
    const int64_t qoff = b * p.q_batch_stride + i * p.ldq + h * T5_HD;
    float q[T5_HD];
    t5_load_row64(p.q, p.q_dtype, qoff, q);
    const int64_t bk = b / p.kv_batch_div;
    const uint8_t* km = p.key_mask != nullptr ? p.key_mask + bk * p.tk : nullptr;
    const float* bias = p.bias != nullptr ? p.bias + static_cast<int64_t>(h) * p.bias_len : nullptr;
    const int jend = p.causal ? min(p.tk, qpos + 1) : p.tk;
    const int64_t kbase = bk * p.kv_batch_stride + h * T5_HD;
    float mx = -INFINITY;
    bool any = false;
    for (int j = lane; j < jend; j += 32) {
      float s = -INFINITY;
      if (km == nullptr || km[j] != 0) {
        float kr[T5_HD];
        t5_load_row64(p.k, p.kv_dtype, kbase + j * p.ldk, kr);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < T5_HD; d += 4) {
          a0 = fmaf(q[d], kr[d], a0);
          a1 = fmaf(q[d + 1], kr[d + 1], a1);
          a2 = fmaf(q[d + 2], kr[d + 2], a2);
          a3 = fmaf(q[d + 3], kr[d + 3], a3);
        }
        s = (a0 + a1) + (a2 + a3);
        if (bias != nullptr) {
          int idx = j - qpos + p.bias_zero;
          idx = idx < 0 ? 0 : (idx >= p.bias_len ? p.bias_len - 1 : idx);
          s += bias[idx];
        }
        any = true;
      }
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    any = __any_sync(0xffffffffu, any);
    // every admissible key masked: the additive finfo.min mask of the reference yields uniform weights over them all
    float sum = 0.f;
    for (int j = lane; j < jend; j += 32) {
      const float e = any ? (sc[j] == -INFINITY ? 0.f : expf(sc[j] - mx)) : 1.f;
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    // P V: lane l accumulates its own keys l, l + 32, ... over all 64 dims (whole value rows, eight independent
    // 16-byte loads in flight), then a butterfly reduce-scatter leaves dims 2l, 2l + 1 of the sum in lane l
    float acc[T5_HD];
#pragma unroll
    for (int d = 0; d < T5_HD; ++d) acc[d] = 0.f;
    for (int j = lane; j < jend; j += 32) {
      const float pj = sc[j];
      if (pj != 0.f) {
        float vr[T5_HD];
        t5_load_row64(p.v, p.kv_dtype, kbase + j * p.ldv, vr);
#pragma unroll
        for (int d = 0; d < T5_HD; ++d) acc[d] = fmaf(pj, vr[d], acc[d]);
      }
    }
    // reduce-scatter over the 32 lanes: after the step with mask m a lane keeps the half of its values that matches
    // its bit m and adds the partner's copy of that half
#pragma unroll
    for (int half = 32, m = 16; m >= 1; half >>= 1, m >>= 1) {
      const bool upper = (lane & m) != 0;
#pragma unroll
      for (int d = 0; d < half; ++d) {
        const float mine = upper ? acc[d + half] : acc[d];
        const float theirs = upper ? acc[d] : acc[d + half];
        acc[d] = mine + __shfl_xor_sync(0xffffffffu, theirs, m);
      }
    }
    // the half kept at mask m = bit log2(m) of the lane picks the 2m-wide block: lane l ends with dims 2l, 2l + 1
    float o0 = acc[0], o1 = acc[1];
    const int pair = lane;
    o0 *= inv, o1 *= inv;
    const int64_t ooff = b * p.o_batch_stride + i * p.ldo;
    const int c = h * T5_HD + 2 * pair;
    const int width = p.num_heads * T5_HD;
    if constexpr (OUT == TSFMX_DT_F32) {
      *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + ooff + c) = make_float2(o0, o1);
    } else if constexpr (OUT == TSFMX_DT_BF16) {
      *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c) = pack_bf16x2(o0, o1);
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + 2 * ooff;  // split rows: [hi(width) | lo(width)]
      uint32_t hi, lo;
      split_bf16x2(o0, o1, hi, lo);
      *reinterpret_cast<uint32_t*>(o + c) = hi;
      *reinterpret_cast<uint32_t*>(o + width + c) = lo;
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------
// Cross-attention of a decode step (one query row per path, bf16 keys / values, no bias, not causal): the kernel that
// bounds Chronos-T5 decoding - every step streams the 513 x 12 heads x (64 k + 64 v) bf16 of every series, 3.2 GB per
// layer at 2048 series.  The general kernel above gives every lane a whole 128-byte key row: a warp-level 16-byte load
// then touches 32 different lines and uses half of each 32-byte sector, the other half being fetched again by the next
// load - ncu: 825 us per launch, 3.9 TB/s, and keeping two rows in flight per lane or putting all heads of a series
// into one block changed nothing (844 us).  Here EIGHT lanes share a key row (8 x 16 B = one full line per key, four
// keys per warp-level load, every sector used once), a lane keeps only its 8 of the 64 dimensions of q, of the
// accumulator and of eight rows in flight (~60 registers), partial dot products meet with three shuffles, and one
// block holds all heads of a series so that it walks the [K | V] rows of its series front to back.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void t5_unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

template <int OUT>
__global__ void __launch_bounds__(384) t5_cross_decode_kernel(const T5AttnParams p) {
  extern __shared__ float s_scores[];  // [warps][tk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int part = lane & 7, kq = lane >> 3;  // 16-byte chunk of the row / which of the four keys of a load
  float* sc = s_scores + warp * p.tk;
  const int b = blockIdx.y;
  const int tk = p.tk;
  const int wpb = blockDim.x >> 5;
  constexpr int U = 8;  // loads in flight per lane: 8 x 4 keys = 32 keys per round
  for (int h = blockIdx.x * wpb + warp; h < p.num_heads; h += gridDim.x * wpb) {
    float q[8];
    {
      const int64_t qoff = b * p.q_batch_stride + h * T5_HD + 8 * part;
      if (p.q_dtype == TSFMX_DT_F32) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.q) + qoff);
        const float4 c = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.q) + qoff + 4);
        q[0] = a.x, q[1] = a.y, q[2] = a.z, q[3] = a.w, q[4] = c.x, q[5] = c.y, q[6] = c.z, q[7] = c.w;
      } else {
        t5_unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.q) + qoff), q);
      }
    }
    const int64_t bk = b / p.kv_batch_div;
    const uint8_t* km = p.key_mask != nullptr ? p.key_mask + bk * tk : nullptr;
    const __nv_bfloat16* kbase = reinterpret_cast<const __nv_bfloat16*>(p.k) + bk * p.kv_batch_stride + h * T5_HD + 8 * part;
    const __nv_bfloat16* vbase = reinterpret_cast<const __nv_bfloat16*>(p.v) + bk * p.kv_batch_stride + h * T5_HD + 8 * part;
    // ---- scores
    float mx = -INFINITY;
    bool any = false;
    for (int j0 = 0; j0 < tk; j0 += 4 * U) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + 4 * u + kq;
        r[u] = j < tk ? ld_stream_u4(kbase + static_cast<int64_t>(j) * p.ldk) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + 4 * u + kq;
        float kf[8];
        t5_unpack8(r[u], kf);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) s = fmaf(q[d], kf[d], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (j < tk) {
          const bool ok = km == nullptr || km[j] != 0;
          s = ok ? s : -INFINITY;
          any = any || ok;
          mx = fmaxf(mx, s);
          if (part == 0) sc[j] = s;
        }
      }
    }
    mx = warp_max(mx);
    any = __any_sync(0xffffffffu, any);
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < tk; j += 32) {
      const float e = any ? (sc[j] == -INFINITY ? 0.f : expf(sc[j] - mx)) : 1.f;  // all masked: uniform (finfo.min mask)
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    // ---- P V: the same ownership; a lane accumulates its 8 dimensions over its quarter of the keys
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < tk; j0 += 4 * U) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + 4 * u + kq;
        r[u] = j < tk ? ld_stream_u4(vbase + static_cast<int64_t>(j) * p.ldv) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + 4 * u + kq;
        const float pj = j < tk ? sc[j] : 0.f;
        float vf[8];
        t5_unpack8(r[u], vf);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(pj, vf[d], acc[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 8);
      acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 16);
      acc[d] *= inv;
    }
    if (kq == 0) {
      const int64_t ooff = b * p.o_batch_stride;
      const int c = h * T5_HD + 8 * part;
      const int width = p.num_heads * T5_HD;
      if constexpr (OUT == TSFMX_DT_F32) {
        float* o = reinterpret_cast<float*>(p.out) + ooff + c;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else if constexpr (OUT == TSFMX_DT_BF16) {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c) =
            make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + 2 * ooff;  // split rows: [hi(width) | lo(width)]
        uint4 hi, lo;
        split_bf16x2(acc[0], acc[1], hi.x, lo.x);
        split_bf16x2(acc[2], acc[3], hi.y, lo.y);
        split_bf16x2(acc[4], acc[5], hi.z, lo.z);
        split_bf16x2(acc[6], acc[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(o + c) = hi;
        *reinterpret_cast<uint4*>(o + width + c) = lo;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core encoder
constexpr int T5_LD = 72;     // padded bf16 row (144 B)
constexpr int T5_WARPS = 8;
constexpr int T5_KC = 4;      // key tiles (of 16) per online-softmax chunk

__device__ __forceinline__ void t5_cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void t5_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void t5_ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void t5_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(T5_WARPS * 32, 1) t5_encoder_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, int seq, int seq_pad, int num_heads, const uint8_t* __restrict__ key_mask,
    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_t5[];
  const int T = seq, TP = seq_pad;  // TP = T rounded up to 64 (a whole number of key chunks)
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_t5);
  __nv_bfloat16* sV = sK + TP * T5_LD;
  __nv_bfloat16* sQ = sV + TP * T5_LD;                                   // [warps][16][LD]
  float* s_bias = reinterpret_cast<float*>(sQ + T5_WARPS * 16 * T5_LD);  // [2T - 1]: index (key - query) + T - 1
  uint8_t* s_valid = reinterpret_cast<uint8_t*>(s_bias + 2 * TP);        // [TP]
  __shared__ int s_any;
  const int b = blockIdx.x / num_heads, h = blockIdx.x - b * num_heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = num_heads * T5_HD;
  const int64_t ld = 3 * static_cast<int64_t>(width);
  const __nv_bfloat16* gbase = qkv + static_cast<int64_t>(b) * T * ld + h * T5_HD;

  if (threadIdx.x == 0) s_any = 0;
  for (int c = threadIdx.x; c < T * 8; c += blockDim.x) {
    const int row = c >> 3, ch = c & 7;
    t5_cp_async_16(sK + row * T5_LD + ch * 8, gbase + row * ld + width + ch * 8);
    t5_cp_async_16(sV + row * T5_LD + ch * 8, gbase + row * ld + 2 * width + ch * 8);
  }
  for (int c = threadIdx.x; c < (TP - T) * 8; c += blockDim.x) {
    const int row = T + (c >> 3), ch = c & 7;
    *reinterpret_cast<uint4*>(sK + row * T5_LD + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sV + row * T5_LD + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = threadIdx.x; i < 2 * T - 1; i += blockDim.x)
    s_bias[i] = bias != nullptr ? __ldg(bias + static_cast<int64_t>(h) * (2 * T - 1) + i) : 0.f;
  __syncthreads();
  {
    int any = 0;
    for (int j = threadIdx.x; j < TP; j += blockDim.x) {
      const uint8_t v = j < T && (key_mask == nullptr || key_mask[static_cast<int64_t>(b) * T + j] != 0) ? 1 : 0;
      s_valid[j] = v;
      any |= v;
    }
    if (any) s_any = 1;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const bool has_key = s_any != 0;
  const int g = lane >> 2, t = lane & 3;
  const int ntq = (T + 15) >> 4;
  const int nchunks = TP / (16 * T5_KC);
  __nv_bfloat16* myQ = sQ + warp * 16 * T5_LD;
  const uint32_t sq_addr = smem_u32(myQ), sk_addr = smem_u32(sK), sv_addr = smem_u32(sV);

  for (int qi = warp; qi < ntq; qi += T5_WARPS) {
    // ---- this tile's 16 query rows -> the warp's own staging slot
    for (int c = lane; c < 16 * 8; c += 32) {
      const int r = c >> 3, ch = c & 7;
      const int row = qi * 16 + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (row < T) v = *reinterpret_cast<const uint4*>(gbase + row * ld + ch * 8);
      *reinterpret_cast<uint4*>(myQ + r * T5_LD + ch * 8) = v;
    }
    __syncwarp();
    uint32_t qf[4][4];
    {
      const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
      const int col = 8 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) t5_ldmatrix_x4(sq_addr + (row * T5_LD + col + 16 * ks) * 2, qf[ks]);
    }
    const int row0 = qi * 16 + g, row1 = row0 + 8;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;

    for (int kc = 0; kc < nchunks; ++kc) {
      float s[T5_KC][2][4];
#pragma unroll
      for (int kj = 0; kj < T5_KC; ++kj) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[kj][nt][e] = 0.f;
        const int key = (kc * T5_KC + kj) * 16 + (lane & 7) + 8 * (lane >> 4);
        const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kb[4];
          t5_ldmatrix_x4(sk_addr + (key * T5_LD + col + 16 * ks) * 2, kb);
          t5_mma(s[kj][0], qf[ks], kb[0], kb[1]);
          t5_mma(s[kj][1], qf[ks], kb[2], kb[3]);
        }
      }
      // bias + mask, chunk row max
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int kj = 0; kj < T5_KC; ++kj)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int key = (kc * T5_KC + kj) * 16 + nt * 8 + 2 * t;
          const uint32_t vv = *reinterpret_cast<const uint16_t*>(s_valid + key);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kk = key + (e & 1);
            const int row = (e & 2) ? row1 : row0;
            const bool ok = has_key ? ((vv >> (8 * (e & 1))) & 0xffu) != 0 : kk < T;
            // rows beyond T read bias out of range: clamp the index, their output is never stored
            int bi = kk - row + T - 1;
            bi = bi < 0 ? 0 : (bi > 2 * T - 2 ? 2 * T - 2 : bi);
            const float v = ok ? (has_key ? s[kj][nt][e] + s_bias[bi] : 0.f) : -INFINITY;
            s[kj][nt][e] = v;
            if (e & 2) cm1 = fmaxf(cm1, v); else cm0 = fmaxf(cm0, v);
          }
        }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float n0 = fmaxf(m0, cm0), n1 = fmaxf(m1, cm1);
      // rescale what has been accumulated so far (exp(-inf - finite) = 0 covers the first chunk)
      const float a0 = n0 == -INFINITY ? 1.f : __expf(m0 - n0), a1 = n1 == -INFINITY ? 1.f : __expf(m1 - n1);
      m0 = n0, m1 = n1;
      l0 *= a0, l1 *= a1;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) o[dt][0] *= a0, o[dt][1] *= a0, o[dt][2] *= a1, o[dt][3] *= a1;
      float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
      for (int kj = 0; kj < T5_KC; ++kj)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float mx = (e & 2) ? m1 : m0;
            const float pv = s[kj][nt][e] == -INFINITY ? 0.f : __expf(s[kj][nt][e] - mx);
            s[kj][nt][e] = pv;
            if (e & 2) cs1 += pv; else cs0 += pv;
          }
      l0 += cs0, l1 += cs1;  // per-thread partial sums; reduced over the quad at the end
#pragma unroll
      for (int kj = 0; kj < T5_KC; ++kj) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[kj][0][0], s[kj][0][1]);
        pa[1] = pack_bf16x2(s[kj][0][2], s[kj][0][3]);
        pa[2] = pack_bf16x2(s[kj][1][0], s[kj][1][1]);
        pa[3] = pack_bf16x2(s[kj][1][2], s[kj][1][3]);
        const int key = (kc * T5_KC + kj) * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int col = 8 * (lane >> 4);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vb[4];
          t5_ldmatrix_x4_trans(sv_addr + (key * T5_LD + col + 16 * dp) * 2, vb);
          t5_mma(o[2 * dp], pa, vb[0], vb[1]);
          t5_mma(o[2 * dp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      *reinterpret_cast<uint32_t*>(myQ + g * T5_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
      *reinterpret_cast<uint32_t*>(myQ + (g + 8) * T5_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
    __syncwarp();
    __nv_bfloat16* obase = out + static_cast<int64_t>(b) * T * width + h * T5_HD;
    for (int c = lane; c < 16 * 8; c += 32) {
      const int r = c >> 3, ch = c & 7;
      const int row = qi * 16 + r;
      if (row < T)
        *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * width + ch * 8) =
            *reinterpret_cast<const uint4*>(myQ + r * T5_LD + ch * 8);
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------
// Next-token choice of a decode step in ONE kernel: ban a token, temperature, top-k, softmax over the survivors,
// inverse-CDF sampling with a caller-supplied uniform number - what upstream's generate(do_sample=True, top_k=50,
// temperature=1.0) does with topk + softmax + multinomial (three eager launches and a [B, V] round trip each).
// top_k = 1 is greedy decoding (the first maximum, like argmax).  One block per row; the k-th largest logit is found
// with a 4-pass radix select on order-preserving keys; ties at the threshold are admitted lowest index first.
// ----------------------------------------------------------------------------------------
constexpr int SAMPLE_THREADS = 256;
constexpr int SAMPLE_MAX_PER_THREAD = 32;  // vocabularies up to 8192

__device__ __forceinline__ uint32_t order_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending float order == ascending unsigned order
}

__global__ void __launch_bounds__(SAMPLE_THREADS) t5_sample_topk_kernel(const float* __restrict__ logits, int vocab,
                                                                        int banned_id, float inv_temperature, int top_k,
                                                                        const float* __restrict__ uniform,
                                                                        int64_t* __restrict__ out) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_want;
  __shared__ float s_red[SAMPLE_THREADS / 32];
  __shared__ float s_scan[SAMPLE_THREADS];
  __shared__ int s_tie_scan[SAMPLE_THREADS];
  __shared__ float s_bcast;
  __shared__ int s_pick;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row = blockIdx.x;
  const float* lr = logits + row * vocab;
  const int per = (vocab + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
  // thread t owns the CONTIGUOUS ids [t * per, (t + 1) * per): prefix sums over threads then follow id order
  float v[SAMPLE_MAX_PER_THREAD];
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
    const int id = tid * per + i;
    v[i] = (i < per && id < vocab && id != banned_id) ? lr[id] * inv_temperature : -INFINITY;
  }
  const int k = (top_k <= 0 || top_k > vocab) ? vocab : top_k;
  // ---- radix select: key of the k-th largest value
  uint32_t prefix = 0, mask = 0;
  uint32_t want = static_cast<uint32_t>(k);  // rank from the top inside the current prefix class
  for (int shift = 24; shift >= 0; shift -= 8) {
    s_hist[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
      if (i < per) {
        const uint32_t key = order_key(v[i]);
        if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (acc + s_hist[bin] >= want) break;
        acc += s_hist[bin];
      }
      s_prefix = prefix | (static_cast<uint32_t>(bin) << shift);
      s_want = want - acc;
    }
    __syncthreads();
    prefix = s_prefix, want = s_want;
    mask |= 0xffu << shift;
    __syncthreads();
  }
  const uint32_t thr_key = prefix;  // key of the k-th largest; `want` = how many values EQUAL to it are admitted
  // ---- row maximum (for the softmax) and admission of ties in id order
  float mx = -INFINITY;
  int ties = 0;
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
    if (i < per) {
      mx = fmaxf(mx, v[i]);
      ties += order_key(v[i]) == thr_key ? 1 : 0;
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  s_tie_scan[tid] = ties;
  __syncthreads();
  if (tid == 0) {
    float m = s_red[0];
    for (int w = 1; w < SAMPLE_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
    s_bcast = m;
    int run = 0;
    for (int t = 0; t < SAMPLE_THREADS; ++t) {  // exclusive scan of the tie counts
      const int c = s_tie_scan[t];
      s_tie_scan[t] = run;
      run += c;
    }
  }
  __syncthreads();
  mx = s_bcast;
  int tie_rank = s_tie_scan[tid];
  // ---- probabilities of the admitted values, thread-local sums
  float p[SAMPLE_MAX_PER_THREAD];
  float local = 0.f;
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
    p[i] = 0.f;
    if (i < per) {
      const uint32_t key = order_key(v[i]);
      bool in = key > thr_key;
      if (key == thr_key) in = tie_rank++ < static_cast<int>(want);
      if (in && v[i] != -INFINITY) p[i] = __expf(v[i] - mx);
      local += p[i];
    }
  }
  s_scan[tid] = local;
  __syncthreads();
  if (tid == 0) {
    float run = 0.f;
    for (int t = 0; t < SAMPLE_THREADS; ++t) {
      const float c = s_scan[t];
      s_scan[t] = run;  // exclusive prefix
      run += c;
    }
    s_bcast = run;
    s_pick = -1;
  }
  __syncthreads();
  const float total = s_bcast;
  const float u = uniform != nullptr ? fminf(fmaxf(uniform[row], 0.f), 0.99999994f) : 0.f;
  const float target = u * total;
  // the thread whose [begin, begin + local) interval holds the target walks its values; the LAST admitted value catches
  // a target that rounding pushed past the total
  const float begin = s_scan[tid];
  if (local > 0.f && target >= begin && (target < begin + local)) {
    float run = begin;
    int pick = -1;
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
      if (i < per && p[i] > 0.f) {
        pick = tid * per + i;
        run += p[i];
        if (target < run) break;
      }
    }
    s_pick = pick;
  }
  __syncthreads();
  if (s_pick < 0) {  // target == total after rounding: the last admitted id
    int last = -1;
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i)
      if (i < per && p[i] > 0.f) last = tid * per + i;
    atomicMax(&s_pick, last);
    __syncthreads();
  }
  if (tid == 0) out[row] = s_pick < 0 ? 0 : s_pick;
}

}  // namespace
}  // namespace tsfmx

namespace tsfmx {
int g_t5_general_attention = 0;  // tune key 5: 1 = the general kernel also for decode-step cross-attention (A/B, tests)
}

using namespace tsfmx;

extern "C" int tsfmx_embed_rows(const int64_t* ids, int64_t rows, int32_t dims, int32_t vocab, const float* table,
                                float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(ids != nullptr && table != nullptr && out != nullptr, "embed_rows: NULL pointer");
  TSFMX_REQUIRE(rows >= 0 && dims > 0 && dims % 4 == 0 && vocab > 0, "embed_rows: bad sizes (dims must be a multiple of 4)");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(table) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "embed_rows: table / out must be 16-byte aligned");
  if (rows == 0) return TSFMX_OK;
  const int64_t blocks = (rows + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  embed_rows_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, stream>>>(ids, rows, dims, vocab, table, out);
  return check_last_launch("embed_rows");
}

extern "C" int tsfmx_t5_attention(const void* q, int32_t q_dtype, int64_t ldq, int64_t q_batch_stride, const void* k,
                                  const void* v, int32_t kv_dtype, int64_t ldk, int64_t ldv, int64_t kv_batch_stride,
                                  int64_t batch, int32_t tq, int32_t tk, int32_t num_heads, int32_t head_dim,
                                  int32_t q_pos0, int32_t causal, const uint8_t* key_mask, const float* bias,
                                  int32_t bias_len, int32_t bias_zero, int32_t out_dtype, void* out, int64_t ldo,
                                  int64_t o_batch_stride, int32_t kv_batch_div, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(q != nullptr && k != nullptr && v != nullptr && out != nullptr, "t5_attention: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && tq > 0 && tk > 0 && num_heads > 0, "t5_attention: bad sizes");
  TSFMX_REQUIRE((q_dtype == TSFMX_DT_F32 || q_dtype == TSFMX_DT_BF16) && (kv_dtype == TSFMX_DT_F32 || kv_dtype == TSFMX_DT_BF16),
                "t5_attention: q / k / v must be f32 or bf16");
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "t5_attention: bad out_dtype");
  TSFMX_REQUIRE(bias == nullptr || bias_len > 0, "t5_attention: bias needs bias_len");
  if (head_dim != T5_HD) {
    set_error("t5_attention: head_dim %d unsupported (T5 uses 64)", head_dim);
    return TSFMX_ERR_UNSUPPORTED;
  }
  TSFMX_REQUIRE(batch < 65536, "t5_attention: batch (%lld) must be below 65536 per call", static_cast<long long>(batch));
  {
    auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
    const int qa = q_dtype == TSFMX_DT_F32 ? 4 : 8, ka = kv_dtype == TSFMX_DT_F32 ? 4 : 8;  // elements per 16 bytes
    TSFMX_REQUIRE(al16(q) && al16(k) && al16(v) && reinterpret_cast<uintptr_t>(out) % 8 == 0 && ldq % qa == 0 &&
                      q_batch_stride % qa == 0 && ldk % ka == 0 && ldv % 2 == 0 && kv_batch_stride % ka == 0 &&
                      ldo % 2 == 0 && o_batch_stride % 2 == 0,
                  "t5_attention: rows must start on 16-byte boundaries");
  }
  if (batch == 0) return TSFMX_OK;
  T5AttnParams p = {};
  p.q = q, p.k = k, p.v = v, p.q_dtype = q_dtype, p.kv_dtype = kv_dtype;
  p.ldq = ldq, p.ldk = ldk, p.ldv = ldv, p.q_batch_stride = q_batch_stride, p.kv_batch_stride = kv_batch_stride;
  p.tq = tq, p.tk = tk, p.num_heads = num_heads, p.q_pos0 = q_pos0, p.causal = causal;
  p.key_mask = key_mask, p.bias = bias, p.bias_len = bias_len, p.bias_zero = bias_zero;
  p.out = out, p.out_dtype = out_dtype, p.ldo = ldo, p.o_batch_stride = o_batch_stride;
  p.kv_batch_div = kv_batch_div < 1 ? 1 : kv_batch_div;
  const int smem = 4 * tk * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) {
    set_error("t5_attention: %d keys need %d bytes of shared memory; unsupported", tk, smem);
    return TSFMX_ERR_UNSUPPORTED;
  }
  const int64_t units = static_cast<int64_t>(tq) * num_heads;
  int gx = static_cast<int>((units + 3) / 4);
  const int cap = num_sms() * 8;
  if (gx > cap) gx = cap;
  const dim3 grid(gx, static_cast<unsigned>(batch));
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("t5_attention: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
    kern<<<grid, 128, smem, stream>>>(p);
    return check_last_launch("t5_attention");
  };
  // a decode step's cross-attention: one query row, bf16 keys / values, no bias, not causal, every row 16-byte aligned
  const bool out16 = reinterpret_cast<uintptr_t>(out) % 16 == 0 && o_batch_stride % 8 == 0;
  if (tq == 1 && kv_dtype == TSFMX_DT_BF16 && !causal && bias == nullptr && ldv % 8 == 0 && tk >= 64 && out16 && !g_t5_general_attention) {
    const int wpb = num_heads <= 12 ? num_heads : 12;  // all heads of a series in one block (T5-base: 12)
    const int smem_x = wpb * tk * static_cast<int>(sizeof(float));
    const dim3 grid_x((num_heads + wpb - 1) / wpb, static_cast<unsigned>(batch));
    auto launch_x = [&](auto kern) -> int {
      if (smem_x > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_x);
        if (e != cudaSuccess) {
          set_error("t5_attention: cudaFuncSetAttribute(%d): %s", smem_x, cudaGetErrorString(e));
          return TSFMX_ERR_CUDA;
        }
      }
      kern<<<grid_x, wpb * 32, smem_x, stream>>>(p);
      return check_last_launch("t5_cross_decode");
    };
    if (smem_x <= 200 * 1024) {
      if (out_dtype == TSFMX_DT_F32) return launch_x(t5_cross_decode_kernel<TSFMX_DT_F32>);
      if (out_dtype == TSFMX_DT_BF16) return launch_x(t5_cross_decode_kernel<TSFMX_DT_BF16>);
      return launch_x(t5_cross_decode_kernel<TSFMX_DT_BF16_SPLIT>);
    }
  }
  if (out_dtype == TSFMX_DT_F32) return launch(t5_attention_kernel<TSFMX_DT_F32>);
  if (out_dtype == TSFMX_DT_BF16) return launch(t5_attention_kernel<TSFMX_DT_BF16>);
  return launch(t5_attention_kernel<TSFMX_DT_BF16_SPLIT>);
}

extern "C" int tsfmx_t5_encoder_attention_mma(const void* qkv, int64_t batch, int32_t seq, int32_t num_heads,
                                              int32_t head_dim, const uint8_t* key_mask, const float* bias,
                                              void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(qkv != nullptr && out != nullptr, "t5_encoder_attention_mma: NULL pointer");
  TSFMX_REQUIRE(batch >= 0 && seq > 0 && num_heads > 0, "t5_encoder_attention_mma: bad sizes");
  TSFMX_REQUIRE(batch * num_heads < (int64_t(1) << 31), "t5_encoder_attention_mma: too many (series, head) pairs");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "t5_encoder_attention_mma: pointers must be 16-byte aligned");
  const int seq_pad = (seq + 63) / 64 * 64;
  const int smem = 2 * seq_pad * T5_LD * 2 + T5_WARPS * 16 * T5_LD * 2 + 2 * seq_pad * 4 + seq_pad;
  if (head_dim != T5_HD || smem > 226 * 1024) {
    set_error("t5_encoder_attention_mma: head_dim %d / seq %d unsupported (head_dim 64, seq <= 704)", head_dim, seq);
    return TSFMX_ERR_UNSUPPORTED;
  }
  if (batch == 0) return TSFMX_OK;
  auto kern = t5_encoder_attention_mma_kernel;
  static int smem_set_dev[64] = {0};  // cudaFuncSetAttribute is per device
  int& smem_set = smem_set_dev[current_device()];
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("t5_encoder_attention_mma: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
    smem_set = smem;
  }
  kern<<<static_cast<int>(batch * num_heads), T5_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), seq, seq_pad, num_heads, key_mask, bias,
      reinterpret_cast<__nv_bfloat16*>(out));
  return check_last_launch("t5_encoder_attention_mma");
}

extern "C" int tsfmx_t5_sample_topk(const float* logits, int64_t rows, int32_t vocab, int32_t banned_id, float temperature,
                                    int32_t top_k, const float* uniform, int64_t* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(logits != nullptr && out != nullptr, "t5_sample_topk: NULL pointer");
  TSFMX_REQUIRE(rows >= 0 && vocab > 0 && vocab <= SAMPLE_THREADS * SAMPLE_MAX_PER_THREAD,
                "t5_sample_topk: vocab %d out of range (<= %d)", vocab, SAMPLE_THREADS * SAMPLE_MAX_PER_THREAD);
  TSFMX_REQUIRE(temperature > 0.f, "t5_sample_topk: temperature must be positive");
  TSFMX_REQUIRE(uniform != nullptr || top_k == 1, "t5_sample_topk: sampling needs one uniform number per row");
  if (rows == 0) return TSFMX_OK;
  t5_sample_topk_kernel<<<static_cast<unsigned>(rows), SAMPLE_THREADS, 0, stream>>>(logits, vocab, banned_id,
                                                                                     1.0f / temperature, top_k, uniform, out);
  return check_last_launch("t5_sample_topk");
}

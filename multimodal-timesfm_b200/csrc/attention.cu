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


// ----------------------------------------------------------------------------------------
// Tensor-core version for the throughput mode (bf16 qkv in, bf16 out).
//
// One warp owns one (series, head).  Its q/k/v rows (N x 80 bf16 each, 160-byte row segments of the
// [B*N, 3*H*hd] qkv matrix) are fetched with 16-byte cp.async into padded shared-memory tiles, q and k
// are conditioned in place in fp32 (RoPE from a per-block cos/sin table, RMSNorm over head_dim,
// per-dim query scale; two lanes per row), and the products S = Q K^T and O = P V run on
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with ldmatrix-fed fragments.  Sequences here have at most
// 64 patches, so a whole score row block lives in registers: no online-softmax rescaling is needed.
// tcgen05 is deliberately not used: its minimum tile is 64 rows and a head has 16 (ctx 512).
// ----------------------------------------------------------------------------------------
constexpr int MMA_HD = 80;
constexpr int MMA_LD = 88;  // padded row (176 B): ldmatrix rows land in distinct bank groups

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>  // number of 16-row tiles: N <= 16 * NT; NT warps share one (series, head), one query tile each
__global__ void __launch_bounds__(NT == 1 ? 256 : 512) timesfm_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, int64_t batch, int num_patches, int num_heads,
    const uint8_t* __restrict__ patch_mask, const int32_t* __restrict__ num_masked,
    const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w, const float* __restrict__ k_ln_w,
    const float* __restrict__ q_scale, float eps, __nv_bfloat16* __restrict__ out) {
  constexpr int ROWS = 16 * NT;
  constexpr int HALF = MMA_HD / 2;          // 40 rotary pairs
  constexpr int TILE = ROWS * MMA_LD;       // elements per q / k / v tile
  extern __shared__ __align__(16) uint8_t smem_attn[];
  float* s_rope = reinterpret_cast<float*>(smem_attn);                    // [2 * ROWS][cos 40 | sin 40 | -sin 40]
  float* s_wq = s_rope + 2 * ROWS * 3 * HALF;                             // [80] q_ln_w * q_scale
  float* s_wk = s_wq + MMA_HD;                                            // [80] k_ln_w
  __nv_bfloat16* s_tiles = reinterpret_cast<__nv_bfloat16*>(s_wk + MMA_HD);
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = num_patches;
  // unit = one (series, head); its NT warps stage and condition the rows together, then warp wq owns query tile wq
  const int units_per_block = warps_per_block / NT;
  const int unit_in_block = warp / NT, wq = warp - unit_in_block * NT;
  const int ulane = wq * 32 + lane;          // lane index inside the unit
  constexpr int UTHREADS = 32 * NT;
  auto unit_sync = [&]() {
    if constexpr (NT == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + unit_in_block), "n"(UTHREADS) : "memory");
  };
  __nv_bfloat16* sQ = s_tiles + unit_in_block * 3 * TILE;
  __nv_bfloat16* sK = sQ + TILE;
  __nv_bfloat16* sV = sK + TILE;

  // ---- per-block tables: cos/sin of (pos * inv_freq) for pos in [-N, N), and the folded norm weights
  for (int i = threadIdx.x; i < 2 * N * HALF; i += blockDim.x) {
    const int p = i / HALF, f = i - p * HALF;
    float sn, cs;
    sincosf(static_cast<float>(p - N) * __ldg(inv_freq + f), &sn, &cs);
    float* row = s_rope + p * 3 * HALF;
    row[f] = cs, row[HALF + f] = sn, row[2 * HALF + f] = -sn;
  }
  for (int i = threadIdx.x; i < MMA_HD; i += blockDim.x) {
    s_wq[i] = __ldg(q_ln_w + i) * __ldg(q_scale + i);
    s_wk[i] = __ldg(k_ln_w + i);
  }
  __syncthreads();

  const int width = num_heads * MMA_HD;
  const int64_t qkv_ld = 3 * static_cast<int64_t>(width);
  const int64_t total = batch * num_heads;
  const int g = lane >> 2, t = lane & 3;
  const int ntk = (N + 15) >> 4;  // 16-row tiles actually populated

  for (int64_t unit = static_cast<int64_t>(blockIdx.x) * units_per_block + unit_in_block; unit < total;
       unit += static_cast<int64_t>(gridDim.x) * units_per_block) {
    const int64_t b = unit / num_heads;
    const int h = static_cast<int>(unit - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;

    // ---- async fetch of the raw q / k / v rows (10 x 16 B per row)
    //      a row is 10 chunks: lane (fr, fch) walks rows fr, fr + RPI, ... with pointer increments only (the
    //      index arithmetic of a flat chunk loop - divisions by N * 10 and 10 - was 38 % of this kernel's instructions)
    const __nv_bfloat16* gbase = qkv + b * N * qkv_ld + h * MMA_HD;
    {
      constexpr int RPI = UTHREADS / 10;  // rows per iteration
      const int fr = ulane / 10, fch = ulane - fr * 10;
      if (fr < RPI) {
        const __nv_bfloat16* gp = gbase + fr * qkv_ld + fch * 8;
        __nv_bfloat16* sp = sQ + fr * MMA_LD + fch * 8;
        for (int row = fr; row < N; row += RPI, gp += RPI * qkv_ld, sp += RPI * MMA_LD) {
          cp_async_16(sp, gp);
          cp_async_16(sp + TILE, gp + width);
          cp_async_16(sp + 2 * TILE, gp + 2 * width);
        }
      }
    }
    // rows beyond N (only when N is not a multiple of 16): zero so the MMAs see finite data
    if (N & 15) for (int c = ulane; c < 3 * (ntk * 16 - N) * 10; c += UTHREADS) {
      const int which = c / ((ntk * 16 - N) * 10);
      const int rem = c - which * ((ntk * 16 - N) * 10);
      const int row = N + rem / 10, ch = rem % 10;
      *reinterpret_cast<uint4*>(sQ + which * TILE + row * MMA_LD + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    // valid-key bitmask (bit j: key j exists and is not padded)
    uint64_t kmask = 0;
    {
      const bool v0 = lane < N && (patch_mask == nullptr || patch_mask[b * N + lane] == 0);
      const bool v1 = lane + 32 < N && (patch_mask == nullptr || patch_mask[b * N + lane + 32] == 0);
      kmask = static_cast<uint64_t>(__ballot_sync(0xffffffffu, v0)) |
              (static_cast<uint64_t>(__ballot_sync(0xffffffffu, v1)) << 32);
    }
    cp_async_wait_all();
    unit_sync();

    // ---- condition q and k in place: lane pair (2r, 2r+1) owns row r; each lane 20 rotary pairs; warp wq takes the
    //      16 rows of its own tile
    {
      const int rt = wq;
      const int r = rt * 16 + (lane >> 1);
      if (rt < ntk) {
        const int hf = lane & 1;
        const bool live = r < N;
        // rotation, sum of squares and scaling on packed fp32 pairs (dims idx, idx + 1 of one tensor): FFMA2 / FMUL2
        // halve the instruction count of what was ~600 of this kernel's ~1 160 warp instructions per (series, head)
        const float* rope = s_rope + (live ? (r - nm + N) : 0) * 3 * HALF + 20 * hf;
        __nv_bfloat16* qrow = sQ + r * MMA_LD + 20 * hf;
        __nv_bfloat16* krow = sK + r * MMA_LD + 20 * hf;
        float2 q1[10], q2[10], k1[10], k2[10];
        float2 qss2 = make_float2(0.f, 0.f), kss2 = make_float2(0.f, 0.f);
        auto unpack = [](uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const uint2 a1 = *reinterpret_cast<const uint2*>(qrow + 4 * i);
          const uint2 a2 = *reinterpret_cast<const uint2*>(qrow + HALF + 4 * i);
          const uint2 c1 = *reinterpret_cast<const uint2*>(krow + 4 * i);
          const uint2 c2 = *reinterpret_cast<const uint2*>(krow + HALF + 4 * i);
          const uint32_t aw1[2] = {a1.x, a1.y}, aw2[2] = {a2.x, a2.y}, cw1[2] = {c1.x, c1.y}, cw2[2] = {c2.x, c2.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * i + e;  // pair index: dims 2j, 2j + 1 of this lane's 20
            const float2 cs = *reinterpret_cast<const float2*>(rope + 2 * j);
            const float2 sn = *reinterpret_cast<const float2*>(rope + HALF + 2 * j);
            const float2 ns = *reinterpret_cast<const float2*>(rope + 2 * HALF + 2 * j);
            const float2 x1 = unpack(aw1[e]), x2 = unpack(aw2[e]), y1 = unpack(cw1[e]), y2 = unpack(cw2[e]);
            q1[j] = fma2(x2, ns, mul2(x1, cs));  // x1 cos - x2 sin
            q2[j] = fma2(x1, sn, mul2(x2, cs));  // x2 cos + x1 sin
            k1[j] = fma2(y2, ns, mul2(y1, cs));
            k2[j] = fma2(y1, sn, mul2(y2, cs));
            qss2 = fma2(q2[j], q2[j], fma2(q1[j], q1[j], qss2));
            kss2 = fma2(k2[j], k2[j], fma2(k1[j], k1[j], kss2));
          }
        }
        float qss = qss2.x + qss2.y, kss = kss2.x + kss2.y;
        qss += __shfl_xor_sync(0xffffffffu, qss, 1);
        kss += __shfl_xor_sync(0xffffffffu, kss, 1);
        const float qrs = rsqrtf(qss * (1.0f / MMA_HD) + eps);
        const float krs = rsqrtf(kss * (1.0f / MMA_HD) + eps);
        const float2 qrs2 = make_float2(qrs, qrs), krs2 = make_float2(krs, krs);
        const float* wq1 = s_wq + 20 * hf;
        const float* wk1 = s_wk + 20 * hf;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          uint32_t o1[2], o2[2], p1[2], p2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * i + e;
            const float2 a = mul2(q1[j], mul2(*reinterpret_cast<const float2*>(wq1 + 2 * j), qrs2));
            const float2 b2 = mul2(q2[j], mul2(*reinterpret_cast<const float2*>(wq1 + HALF + 2 * j), qrs2));
            const float2 c = mul2(k1[j], mul2(*reinterpret_cast<const float2*>(wk1 + 2 * j), krs2));
            const float2 d = mul2(k2[j], mul2(*reinterpret_cast<const float2*>(wk1 + HALF + 2 * j), krs2));
            o1[e] = pack_bf16x2(a.x, a.y), o2[e] = pack_bf16x2(b2.x, b2.y);
            p1[e] = pack_bf16x2(c.x, c.y), p2[e] = pack_bf16x2(d.x, d.y);
          }
          if (live) {
            *reinterpret_cast<uint2*>(qrow + 4 * i) = make_uint2(o1[0], o1[1]);
            *reinterpret_cast<uint2*>(qrow + HALF + 4 * i) = make_uint2(o2[0], o2[1]);
            *reinterpret_cast<uint2*>(krow + 4 * i) = make_uint2(p1[0], p1[1]);
            *reinterpret_cast<uint2*>(krow + HALF + 4 * i) = make_uint2(p2[0], p2[1]);
          }
        }
      }
    }
    unit_sync();  // every warp of the unit needs all conditioned k rows

    // ---- this warp's 16-query tile: S = Q K^T (tensor cores) -> masked softmax (registers) -> O = P V
    const uint32_t sq_addr = smem_u32(sQ), sk_addr = smem_u32(sK), sv_addr = smem_u32(sV);
    {
      const int qi = wq;
      if (qi < ntk) {
        uint32_t qa[5][4];
        {
          const int row = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int col = 8 * (lane >> 4);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) ldmatrix_x4(sq_addr + (row * MMA_LD + col + 16 * ks) * 2, qa[ks]);
        }
        const int row0 = qi * 16 + g, row1 = row0 + 8;
        // a row whose keys are all masked attends uniformly to ALL N keys (additive finfo.min mask)
        const bool dead0 = (kmask & ((2ull << row0) - 1ull)) == 0ull;
        const bool dead1 = (kmask & ((2ull << row1) - 1ull)) == 0ull;
        const bool any_dead = __any_sync(0xffffffffu, (dead0 && row0 < N) || (dead1 && row1 < N));
        const int kj_end = any_dead ? ntk : qi + 1;

        float s[NT][2][4];
#pragma unroll
        for (int kj = 0; kj < NT; ++kj) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[kj][nt][e] = 0.f;
          if (kj <= qi) {
            const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
            const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
              uint32_t kb[4];
              ldmatrix_x4(sk_addr + (key * MMA_LD + col + 16 * ks) * 2, kb);
              mma_bf16_16816(s[kj][0], qa[ks], kb[0], kb[1]);
              mma_bf16_16816(s[kj][1], qa[ks], kb[2], kb[3]);
            }
          }
        }
        // mask + row max
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int kj = 0; kj < NT; ++kj)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int key = kj * 16 + nt * 8 + 2 * t + (e & 1);
              const int row = (e & 2) ? row1 : row0;
              const bool ok = kj <= qi && key <= row && ((kmask >> key) & 1ull);
              const float v = ok ? s[kj][nt][e] : -INFINITY;
              s[kj][nt][e] = v;
              if (e & 2) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
            }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
        const float ml0 = mx0 * LOG2E_F, ml1 = mx1 * LOG2E_F;  // finite unless the row is dead (handled below)
#pragma unroll
        for (int kj = 0; kj < NT; ++kj)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int key = kj * 16 + nt * 8 + 2 * t + (e & 1);
              const bool dead = (e & 2) ? dead1 : dead0;
              float p;
              if (dead) p = key < N ? 1.f : 0.f;
              else p = exp_sub(s[kj][nt][e], (e & 2) ? ml1 : ml0);  // masked scores are -inf -> 0
              s[kj][nt][e] = p;
              if (e & 2) sum1 += p; else sum0 += p;
            }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

        float o[10][4];
#pragma unroll
        for (int dt = 0; dt < 10; ++dt)
#pragma unroll
          for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
        for (int kj = 0; kj < NT; ++kj) {
          if (kj < kj_end) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[kj][0][0], s[kj][0][1]);
            pa[1] = pack_bf16x2(s[kj][0][2], s[kj][0][3]);
            pa[2] = pack_bf16x2(s[kj][1][0], s[kj][1][1]);
            pa[3] = pack_bf16x2(s[kj][1][2], s[kj][1][3]);
            const int key = kj * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int col = 8 * (lane >> 4);
#pragma unroll
            for (int dp = 0; dp < 5; ++dp) {
              uint32_t vb[4];
              ldmatrix_x4_trans(sv_addr + (key * MMA_LD + col + 16 * dp) * 2, vb);
              mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
              mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
            }
          }
        }
        // the Q rows of this tile are dead now: stage the normalised output there for coalesced stores
        __syncwarp();
#pragma unroll
        for (int dt = 0; dt < 10; ++dt) {
          *reinterpret_cast<uint32_t*>(sQ + row0 * MMA_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
          *reinterpret_cast<uint32_t*>(sQ + row1 * MMA_LD + dt * 8 + 2 * t) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
      }
    }
    __syncwarp();
    // each warp stores the 16 rows it produced (staged in its own Q rows)
    __nv_bfloat16* obase = out + b * N * width + h * MMA_HD;
    {
      const int sr = lane / 10, sch = lane - sr * 10;  // three rows of 10 chunks per pass, two lanes idle
      if (sr < 3) {
        const int row_first = wq * 16 + sr;
        const int row_end = min(N, wq * 16 + 16);
        __nv_bfloat16* gp = obase + static_cast<int64_t>(row_first) * width + sch * 8;
        const __nv_bfloat16* sp = sQ + row_first * MMA_LD + sch * 8;
        for (int row = row_first; row < row_end; row += 3, gp += 3 * width, sp += 3 * MMA_LD)
          *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
      }
    }
    unit_sync();  // the next unit's cp.async must not overwrite k / v rows another warp still reads
  }
}

template <int NT>
int launch_attention_mma(const void* qkv, int64_t batch, int N, int H, const uint8_t* pm, const int32_t* nm,
                         const float* inv_freq, const float* qw, const float* kw, const float* qs, float eps, void* out,
                         cudaStream_t stream) {
  constexpr int ROWS = 16 * NT;
  constexpr int per_unit = 3 * ROWS * MMA_LD * 2;
  constexpr int fixed = 2 * ROWS * 3 * 40 * 4 + 2 * MMA_HD * 4;  // rope table (cos | sin | -sin) + folded norm weights
  int upb = (200 * 1024 - fixed) / per_unit;                 // units ((series, head) pairs) per block
  const int max_units = NT == 1 ? 8 : 512 / (32 * NT);       // NT warps per unit, <= 512 threads per block
  if (upb > max_units) upb = max_units;
  const int wpb = upb * NT;
  const int smem = fixed + upb * per_unit;
  auto kern = timesfm_attention_mma_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("timesfm_attention: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
  }
  const int64_t total = batch * H;
  const int64_t blocks = (total + upb - 1) / upb;
  const int per_sm = (220 * 1024) / smem > 0 ? (220 * 1024) / smem : 1;
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  kern<<<grid, wpb * 32, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), batch, N, H, pm, nm, inv_freq, qw,
                                          kw, qs, eps, reinterpret_cast<__nv_bfloat16*>(out));
  return check_last_launch("timesfm_attention_mma");
}

template <int HD, int QKV_BF16>
int launch_attention(const void* qkv, int64_t batch, int N, int H, const uint8_t* pm, const int32_t* nm,
                     const float* inv_freq, const float* qw, const float* kw, const float* qs, float eps, int out_dtype,
                     void* out, cudaStream_t stream) {
  const int per_warp_bytes = (3 * N * (HD + 1) + N) * 4;
  int wpb = (96 * 1024) / per_warp_bytes;
  if (wpb > 4) wpb = 4;
  if (wpb < 1 && per_warp_bytes <= 224 * 1024) wpb = 1;  // long contexts (up to 224 patches): one warp per block
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

extern "C" int tsfmx_attention_force_simt(int on) {
  g_force_simt_attention = on ? 1 : 0;
  return TSFMX_OK;
}

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
  const bool aligned = reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  if (qkv_dtype == TSFMX_DT_BF16 && out_dtype == TSFMX_DT_BF16 && aligned && num_patches <= 64 && !g_force_simt_attention) {
    if (num_patches <= 16)
      return launch_attention_mma<1>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w,
                                     k_ln_w, q_scale, eps, out, stream);
    if (num_patches <= 32)
      return launch_attention_mma<2>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w,
                                     k_ln_w, q_scale, eps, out, stream);
    return launch_attention_mma<4>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w,
                                   k_ln_w, q_scale, eps, out, stream);
  }
  if (qkv_dtype == TSFMX_DT_BF16)
    return launch_attention<80, 1>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w, k_ln_w,
                                   q_scale, eps, out_dtype, out, stream);
  return launch_attention<80, 0>(qkv, batch, num_patches, num_heads, patch_mask, num_masked, inv_freq, q_ln_w, k_ln_w,
                                 q_scale, eps, out_dtype, out, stream);
}

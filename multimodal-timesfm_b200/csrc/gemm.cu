// tcgen05 / TMEM GEMM fed by TMA — the dense workhorse of the forecast path.
//
//   D[M, N] = epilogue( sum_s A_s[M, K_s] * B_s[N, K_s]^T )
//
// A_s, B_s are bf16, K-major (row-major [rows, K]); nn.Linear.weight is already
// [N, K] K-major so no operand is ever transposed.  Accumulation is fp32 in
// tensor memory.  The K loop walks a list of "segments": that one mechanism
// gives (i) ResidualBlock fusion (hidden path + residual path accumulate into
// the same TMEM tile; reference ResidualBlock = output_layer(act(hidden)) +
// residual_layer(x)) and (ii) the bf16x3 parity mode (hi*hi + hi*lo + lo*hi of
// split operands) without a second kernel.
//
// Kernel organisation (persistent, warp-specialised, one CTA or one CTA pair per SM):
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D tiles, 128B swizzle, STAGES-deep ring
//   warp 1    MMA issuer:   one elected thread issues tcgen05.mma (UMMA 128xBNx16, or
//                           256xBNx16 with cta_group::2), tcgen05.commit frees smem slots
//   warp 2    TMEM allocator (2 x BN fp32 columns: the epilogue of tile i overlaps the
//                           MMAs of tile i+1)
//   warps 4-11 epilogue:    tcgen05.ld 32x32b -> registers -> bias/act/affine/residual ->
//                           vectorised global stores (f32 / bf16 / split bf16)
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace tsfmx {

namespace {

constexpr int BM = 128;      // rows per CTA tile (TMEM lanes)
constexpr int BK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;   // K of one tcgen05.mma kind::f16
constexpr int MAX_XSEG = 6;  // expanded segments (2 logical x 3 split products)
constexpr int NUM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;     // two warps per TMEM lane quarter, each owning half of the tile's columns

struct XSeg {
  int32_t a_idx, b_idx;    // which tensor map
  int32_t a_koff, b_koff;  // element offset along K inside that tensor
  int32_t nkb;             // number of 64-wide k-blocks
};

struct alignas(64) GemmParams {
  CUtensorMap tma_a[2];
  CUtensorMap tma_b[2];
  XSeg xseg[MAX_XSEG];
  int32_t num_xseg;
  int32_t n;
  int64_t m;
  int32_t tiles_m, tiles_n;  // tiles_m counts BM*CG-row tiles
  int32_t act, d_dtype;
  const float* bias;
  const float* row_scale;
  const float* row_shift;
  const float* residual;
  int64_t ldr;
  void* d;
  int64_t ldd;
  int32_t n_store, split_off;
  int32_t vec_ok;  // all pointers / leading dims allow 16-byte vector access
  int32_t aux_vec_ok, pre_vec_ok;  // same for the saved-activation input / pre-activation output of the training path
  int32_t split_k;  // > 1: every (m, n) tile is computed by split_k tiles over disjoint k-block ranges that add their
                    // fp32 partial sums into a zero-filled D with vector reductions (weight gradients: K = tokens)
  const void* aux;  // *_GRAD activations: the saved pre-activation (SiLU) / activation output (ReLU)
  int64_t ld_aux;
  int32_t aux_dtype;
  int32_t pre_dtype;
  void* pre_act;    // optional second output: acc + bias, before the activation (saved for the backward pass)
  int64_t ld_pre;
};

template <int BN, int CG>
struct Cfg {
  static constexpr int B_ROWS = BN / CG;  // B rows held by one CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(2 * BN <= 512, "two accumulator stages must fit in TMEM");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
};

__device__ __forceinline__ float load_aux(const void* base, int dtype, int64_t idx) {
  return dtype == TSFMX_DT_F32 ? reinterpret_cast<const float*>(base)[idx]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

// 32 consecutive saved activations of one row (the *_GRAD epilogues): 16-byte loads when the layout allows.  The
// element-wise version cost a warp 32 load instructions of 32 scattered sectors each and made the SiLU' dgrad GEMM
// four times slower than its mainloop.
__device__ __forceinline__ void load_aux32(const void* base, int dtype, int64_t idx, bool vec, int valid, float (&u)[32]) {
  if (vec && valid == 32) {
    if (dtype == TSFMX_DT_F32) {
      const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = p4[i];
        u[4 * i + 0] = t.x, u[4 * i + 1] = t.y, u[4 * i + 2] = t.z, u[4 * i + 3] = t.w;
      }
    } else {
      const uint4* p4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 t = p4[i];
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          u[8 * i + 2 * k + 0] = __uint_as_float(w[k] << 16);
          u[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) u[i] = i < valid ? load_aux(base, dtype, idx + i) : 0.0f;
  }
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// sigmoid of four values with FIVE special-function ops instead of eight: one shared reciprocal of the product of the
// four denominators (1 + e^-x), the individual reciprocals recovered with multiplies.  The SFU pipe (16 ops / clock /
// SM) was the limit of the SiLU epilogue: 2 x 32768 elements per 128 x 256 tile = 4096 cycles next to ~5100 cycles of
// MMA, which made ff0 23 % slower than the same-size plain GEMM (ncu: 187 us against 152 us, tensor pipe 59 % active).
// Inputs are clamped at -20 (sigmoid 2e-9) so that the product of four denominators stays below 6e34.
__device__ __forceinline__ void sigmoid4(const float (&x)[4], float (&s)[4]) {
  float d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = 1.0f + ex2_approx(fmaxf(x[i], -20.0f) * -1.4426950408889634f);
  const float d01 = d[0] * d[1], d23 = d[2] * d[3];
  const float r = rcp_approx(d01 * d23);
  const float r01 = r * d23, r23 = r * d01;
  s[0] = r01 * d[1], s[1] = r01 * d[0], s[2] = r23 * d[3], s[3] = r23 * d[2];
}

// Epilogue for 32 consecutive columns [col0, col0+32) of one output row.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int64_t row,
                                               int col0) {
  const int n_store = p.n_store;
  if (col0 >= n_store) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

  const bool full = (col0 + 32 <= n_store) && p.vec_ok;
  if (p.bias != nullptr) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x, v[4 * i + 1] += b.y, v[4 * i + 2] += b.z, v[4 * i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_store) v[i] += __ldg(p.bias + col0 + i);
    }
  }
  if (p.pre_act != nullptr) {
    // training forward: keep the pre-activation for the backward pass
    if (col0 + 32 <= n_store && p.pre_vec_ok) {
      if (p.pre_dtype == TSFMX_DT_F32) {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.pre_act) + row * p.ld_pre + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.pre_act) + row * p.ld_pre + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[i] = make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                             pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (col0 + i < n_store) {
          const int64_t o = row * p.ld_pre + col0 + i;
          if (p.pre_dtype == TSFMX_DT_F32) reinterpret_cast<float*>(p.pre_act)[o] = v[i];
          else reinterpret_cast<__nv_bfloat16*>(p.pre_act)[o] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  }
  if (p.act == TSFMX_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float x[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
      float sg[4];
      sigmoid4(x, sg);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[i + k] *= sg[k];
    }
  } else if (p.act == TSFMX_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (p.act == TSFMX_ACT_SILU_GRAD || p.act == TSFMX_ACT_RELU_GRAD) {
    float u[32];
    const int valid = n_store - col0 < 32 ? n_store - col0 : 32;
    load_aux32(p.aux, p.aux_dtype, row * p.ld_aux + col0, p.aux_vec_ok != 0, valid, u);
    if (p.act == TSFMX_ACT_SILU_GRAD) {
      // v = dL/d silu(u)  ->  dL/du = v * sigmoid(u) * (1 + u * (1 - sigmoid(u)))
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float x[4] = {u[i], u[i + 1], u[i + 2], u[i + 3]};
        float sg[4];
        sigmoid4(x, sg);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[i + k] *= sg[k] * (1.0f + u[i + k] * (1.0f - sg[k]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = u[i] > 0.0f ? v[i] : 0.0f;
    }
  }
  if (p.row_scale != nullptr) {
    const float s = __ldg(p.row_scale + row);
    const float t = p.row_shift != nullptr ? __ldg(p.row_shift + row) : 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] * s + t;
  }
  if (p.residual != nullptr) {
    const float* rp = p.residual + row * p.ldr + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(rp + 4 * i);
        v[4 * i + 0] += x.x, v[4 * i + 1] += x.y, v[4 * i + 2] += x.z, v[4 * i + 3] += x.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_store) v[i] += rp[i];
    }
  }

  if (p.split_k > 1) {  // fp32 D, 16-byte aligned rows (checked by the launcher): add this k-range's partial sums
    float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (col0 + 4 * i + 4 <= n_store) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                     "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                     : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (col0 + 4 * i + k < n_store) atomicAdd(dp + 4 * i + k, v[4 * i + k]);
      }
    }
    return;
  }
  if (p.d_dtype == TSFMX_DT_F32) {
    float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dp + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_store) dp[i] = v[i];
    }
  } else if (p.d_dtype == TSFMX_DT_BF16) {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q;
        q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        *reinterpret_cast<uint4*>(dp + 8 * i) = q;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_store) dp[i] = __float2bfloat16_rn(v[i]);
    }
  } else {  // TSFMX_DT_BF16_SPLIT
    __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col0;
    __nv_bfloat16* lp = hp + p.split_off;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 qh, ql;
        split_bf16x2(v[8 * i + 0], v[8 * i + 1], qh.x, ql.x);
        split_bf16x2(v[8 * i + 2], v[8 * i + 3], qh.y, ql.y);
        split_bf16x2(v[8 * i + 4], v[8 * i + 5], qh.z, ql.z);
        split_bf16x2(v[8 * i + 6], v[8 * i + 7], qh.w, ql.w);
        *reinterpret_cast<uint4*>(hp + 8 * i) = qh;
        *reinterpret_cast<uint4*>(lp + 8 * i) = ql;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < n_store) split_bf16(v[i], hp[i], lp[i]);
    }
  }
}

// TOKEN_MAJOR (weight gradients dW = dY^T X, K = tokens): both operands are read as they lie in HBM, [tokens, features]
// row-major, i.e. MN-major for this product.  One stage still holds 128 x 64 elements per operand, but as 64-feature
// chunks of [64 tokens][128 B] (one TMA box each, 128B swizzle): the UMMA descriptors then walk MN in steps of one
// chunk (LBO = 8 KB) and K in 8-token groups of 1 KB (SBO); one UMMA_K = 16 tokens = 2 KB.  Tokens beyond the last row
// are zero-filled by TMA, so K needs no padding - and nothing is transposed through HBM any more.
template <int BN, int CG, bool TOKEN_MAJOR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, CG>;
  constexpr int CHUNK_BYTES = 64 * BK * 2;  // one 64-feature x 64-token box
  constexpr int STAGES = C::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte aligned bases (same offset in every CTA).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES] TMA -> MMA   (leader CTA's are used when CG == 2)
  uint64_t* empty_bar = bars + STAGES;          // [STAGES] MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;  // [2] MMA -> epilogue
  uint64_t* tmem_empty_bar = tmem_full_bar + 2; // [2] epilogue -> MMA  (leader's when CG == 2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a[0]);
    tma_prefetch_desc(&p.tma_b[0]);
    if (p.num_xseg > 1) {
      tma_prefetch_desc(&p.tma_a[1]);
      tma_prefetch_desc(&p.tma_b[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);  // one arrive(+expect_tx) per producing CTA
      mbar_init(&empty_bar[i], 1);  // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);        // one tcgen05.commit
      mbar_init(&tmem_empty_bar[i], EPI_WARPS * CG);  // one arrive per epilogue warp (of every CTA of the pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.tiles_m * p.tiles_n * p.split_k;
  const int first_tile = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int tile_step = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int mn = tile / p.split_k, ks = tile - mn * p.split_k;
        const int m_blk = mn / p.tiles_n, n_blk = mn % p.tiles_n;
        const int32_t a_row = (m_blk * CG + static_cast<int>(cta_rank)) * BM;
        const int32_t b_row = n_blk * BN + static_cast<int>(cta_rank) * C::B_ROWS;
        for (int s = 0; s < p.num_xseg; ++s) {
          const XSeg sg = p.xseg[s];
          const CUtensorMap* ta = &p.tma_a[sg.a_idx];
          const CUtensorMap* tb = &p.tma_b[sg.b_idx];
          const int kb_end = (ks + 1) * sg.nkb / p.split_k;
          for (int kb = ks * sg.nkb / p.split_k; kb < kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            void* sa = smem_a + stage * C::A_BYTES;
            void* sb = smem_b + stage * C::B_BYTES;
            if constexpr (TOKEN_MAJOR) {
              static_assert(!TOKEN_MAJOR || CG == 2, "token-major operands: CTA pairs only");
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * 2);
              else        mbar_arrive_cluster(&full_bar[stage], 0);
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)
                tma_load_2d_pair(static_cast<uint8_t*>(sa) + c * CHUNK_BYTES, ta, &full_bar[stage], a_row + 64 * c, kb * BK);
#pragma unroll
              for (int c = 0; c < C::B_ROWS / 64; ++c)
                tma_load_2d_pair(static_cast<uint8_t*>(sb) + c * CHUNK_BYTES, tb, &full_bar[stage], b_row + 64 * c, kb * BK);
            } else if constexpr (CG == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
              tma_load_2d(sa, ta, &full_bar[stage], sg.a_koff + kb * BK, a_row);
              tma_load_2d(sb, tb, &full_bar[stage], sg.b_koff + kb * BK, b_row);
            } else {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * 2);
              else        mbar_arrive_cluster(&full_bar[stage], 0);
              tma_load_2d_pair(sa, ta, &full_bar[stage], sg.a_koff + kb * BK, a_row);
              tma_load_2d_pair(sb, tb, &full_bar[stage], sg.b_koff + kb * BK, b_row);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN) | (TOKEN_MAJOR ? (1u << 15) | (1u << 16) : 0u);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
        const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t accumulate = 0;
        const int ks = tile % p.split_k;
        for (int s = 0; s < p.num_xseg; ++s) {
          const int nkb = (ks + 1) * p.xseg[s].nkb / p.split_k - ks * p.xseg[s].nkb / p.split_k;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if constexpr (TOKEN_MAJOR) {
              constexpr uint32_t SBO = 8 * 128;  // 8 tokens x 128 B
              const uint64_t da = make_umma_desc_mn_sw128(smem_u32(smem_a + stage * C::A_BYTES), CHUNK_BYTES, SBO);
              const uint64_t db = make_umma_desc_mn_sw128(smem_u32(smem_b + stage * C::B_BYTES), CHUNK_BYTES, SBO);
              constexpr uint32_t kstep = (UMMA_K / 8) * SBO >> 4;  // 16 tokens = two 8-token groups (descriptor units of 16 B)
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                umma_bf16<CG>(tmem_d, da + kstep * k, db + kstep * k, idesc, accumulate);
                accumulate = 1;
              }
            } else {
              const uint64_t da = make_umma_desc_sw128(smem_u32(smem_a + stage * C::A_BYTES));
              const uint64_t db = make_umma_desc_sw128(smem_u32(smem_b + stage * C::B_BYTES));
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                umma_bf16<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            }
            if constexpr (CG == 1) umma_commit(&empty_bar[stage]);
            else                   umma_commit_pair(&empty_bar[stage], 0b11);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if constexpr (CG == 1) umma_commit(&tmem_full_bar[as]);
        else                   umma_commit_pair(&tmem_full_bar[as], 0b11);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int chalf = (warp - EPI_WARP0) >> 2;   // which half of the tile's columns this warp drains
    uint32_t iter = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
      const int mn = tile / p.split_k;
      const int m_blk = mn / p.tiles_n, n_blk = mn % p.tiles_n;
      const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(m_blk * CG + static_cast<int>(cta_rank)) * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
      const int col_base = n_blk * BN;
#pragma unroll 1
      for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);  // warp-collective: issued even for out-of-range rows
        tmem_ld_wait();
        if (row < p.m) epilogue_chunk(p, r, row, col_base + c * 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tmem_empty_bar[as]);
        else                   mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
  }
}


// =====================================================================================================
// GEMM with the row-norm / residual junction of a transformer layer fused into its epilogue:
//
//     a  = A W^T                               (K-loop as above; a never leaves the chip)
//     y  = RMSNorm(a) * w_post + x             (w_post NULL: y = a + x, the pre-norm form of Chronos-2)
//     yn = RMSNorm(y) * w_next                 (w_next NULL: yn = y)
//
// i.e. out-proj / ff1 of a TimesFM 2.5 layer together with `post_ln(.) + x` and the next `pre_ln`
// (HF twin modeling_timesfm2_5.py:378-388) — the work of norm_residual_norm_kernel without the HBM
// round trip of `a` and without a separate launch.  RMSNorm needs statistics over the whole row
// (N = 1280) while one CTA's TMEM holds 256 columns, so C = N / 256 CTAs form a thread-block cluster
// that owns a full 128-row x N output panel: every CTA runs its own 128x256 tcgen05 tile, the A tile is
// fetched ONCE per cluster (each issuer CTA multicasts a row slice to all C CTAs) and the per-row sums of
// squares are exchanged through distributed shared memory with st.async + mbarrier complete_tx (no cluster
// barrier, no fence).  y is written back into the accumulator's TMEM columns between the two passes.
// =====================================================================================================
constexpr int RN_BN = 256;
constexpr int RN_MAXC = 5;
constexpr int RN_STAGES = 3;
constexpr int RN_XCHUNK = 32;                         // fp32 columns per staged x / y chunk (128 B: one swizzle row)
constexpr int RN_XBUF_BYTES = BM * RN_XCHUNK * 4;     // 16 KB
constexpr int RN_XSLOTS = 2;                          // chunk buffers per column half

struct alignas(64) RownormParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_x;    // fp32 [m, n], box 32 x 128, 128B swizzle (load)
  CUtensorMap tma_y;    // same geometry over y (store); y may alias x
  XSeg xseg[3];
  int32_t num_xseg;
  int32_t n;
  int64_t m;
  int32_t tiles_m;
  int32_t cluster;      // C = n / 256
  int32_t a_issuers;    // CTAs that fetch a slice of the A tile (4 when C >= 4, else 2)
  int32_t yn_dtype;
  const float* w_post;
  const float* w_next;
  void* yn;
  float eps;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0,
                                                  int32_t c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Second design of the epilogue data path (the first one read x and wrote y with one global row segment per
// thread and was latency bound, 2x slower than GEMM + norm kernel): x arrives and y leaves through 128-row x
// 32-column fp32 chunks in 128B-swizzled shared memory, moved by TMA - loads issued one chunk ahead by an otherwise
// idle control warp, stores asynchronous (bulk groups) - so the epilogue warps only touch TMEM and shared memory
// (conflict-free 16-byte accesses) and the HBM traffic of the junction overlaps the MMAs of the next tile.
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_rownorm_tcgen05_kernel(const __grid_constant__ RownormParams p) {
  constexpr int STAGES = RN_STAGES;
  constexpr int A_BYTES = BM * BK * 2;        // 16 KB
  constexpr int B_BYTES = RN_BN * BK * 2;     // 32 KB
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int RED_FLOATS = 2 * 2 * RN_MAXC * BM;  // [acc stage][round][src CTA][row]

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* xbuf = smem + STAGES * STAGE_BYTES;                       // [2 halves][RN_XSLOTS][16 KB], 1024-aligned
  float* red = reinterpret_cast<float*>(xbuf + 2 * RN_XSLOTS * RN_XBUF_BYTES);
  float* pairbuf = red + RED_FLOATS;                                 // [2 rounds][BM]: column half 1 -> half 0
  float* wbuf = pairbuf + 2 * BM;                                    // [2][RN_BN]: this CTA's w_post / w_next slices
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + 2 * RN_BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* red_bar = tmem_empty_bar + 2;   // [acc stage][round]
  uint64_t* xfull_bar = red_bar + 4;        // [half][slot]
  uint64_t* xempty_bar = xfull_bar + 2 * RN_XSLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty_bar + 2 * RN_XSLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.cluster;
  const uint32_t rank = cluster_ctarank();
  const uint16_t all_mask = static_cast<uint16_t>((1u << C) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    tma_prefetch_desc(&p.tma_x);
    tma_prefetch_desc(&p.tma_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);                          // own arrive(+expect_tx); bytes come from C issuers
      mbar_init(&empty_bar[i], static_cast<uint32_t>(C));  // every CTA's MMA must be done with the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_WARPS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&red_bar[i], 1);
    for (int i = 0; i < 2 * RN_XSLOTS; ++i) {
      mbar_init(&xfull_bar[i], 1);
      mbar_init(&xempty_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  // norm weights of this CTA's 256 columns: read once, then broadcast from shared memory (the L1 left next to
  // ~225 KB of shared memory does not keep them, and an L2 round trip per chunk stalled the epilogue)
  for (int i = threadIdx.x; i < 2 * RN_BN; i += blockDim.x) {
    const float* src = i < RN_BN ? p.w_post : p.w_next;
    wbuf[i] = src != nullptr ? __ldg(src + cluster_ctarank() * RN_BN + (i % RN_BN)) : 1.0f;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first_tile = blockIdx.x / C;
  const int tile_step = gridDim.x / C;
  const int rows_per_issuer = BM / p.a_issuers;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (A multicast slices, own B tile)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = first_tile; tile < p.tiles_m; tile += tile_step) {
        const int32_t a_row = tile * BM + static_cast<int>(rank) * rows_per_issuer;
        const int32_t b_row = static_cast<int>(rank) * RN_BN;
        for (int s = 0; s < p.num_xseg; ++s) {
          const XSeg sg = p.xseg[s];
          for (int kb = 0; kb < sg.nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(smem_b + stage * B_BYTES, &p.tma_b, &full_bar[stage], sg.b_koff + kb * BK, b_row);
            if (static_cast<int>(rank) < p.a_issuers)
              tma_load_2d_mcast(smem_a + stage * A_BYTES + static_cast<int>(rank) * rows_per_issuer * BK * 2, &p.tma_a,
                                &full_bar[stage], sg.a_koff + kb * BK, a_row, all_mask);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (every CTA: cta_group::1)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, RN_BN);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int tile = first_tile; tile < p.tiles_m; tile += tile_step, ++iter) {
        const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * RN_BN;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.num_xseg; ++s) {
          const int nkb = p.xseg[s].nkb;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = make_umma_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
            const uint64_t db = make_umma_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_bf16<1>(tmem_d, da + 2 * k, db + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit_mcast(&empty_bar[stage], all_mask);  // the A slice we multicast lives in every CTA
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[as]);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ x producer: residual chunks one step ahead
    if (lane == 0) {
      uint32_t n = 0;  // chunk counter (same sequence in both column halves)
      for (int tile = first_tile; tile < p.tiles_m; tile += tile_step) {
        for (int c = 0; c < 4; ++c, ++n) {
          const uint32_t slot = n % RN_XSLOTS, phase = (n / RN_XSLOTS) & 1;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int bi = hh * RN_XSLOTS + static_cast<int>(slot);
            mbar_wait(&xempty_bar[bi], phase ^ 1);
            mbar_arrive_expect_tx(&xfull_bar[bi], RN_XBUF_BYTES);
            tma_load_2d(xbuf + bi * RN_XBUF_BYTES, &p.tma_x, &xfull_bar[bi],
                        static_cast<int>(rank) * RN_BN + hh * 128 + c * RN_XCHUNK, tile * BM);
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------ epilogue: norm / residual / norm
    const int q = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const int row_in_tile = q * 32 + lane;
    const bool two_norms = p.w_post != nullptr;  // pass 1 (statistics of a) only exists in the post-norm form
    const bool norm_out = p.w_next != nullptr && p.yn != nullptr;
    const float inv_n = 1.0f / static_cast<float>(p.n);
    const uint32_t red_bytes = static_cast<uint32_t>(C) * BM * 4;
    const bool store_issuer = q == 0 && lane == 0;  // one thread per column half issues the y stores
    // 16-byte unit u of this thread's row sits at unit (u ^ (row & 7)) of the 128-byte swizzled line
    const uint32_t xrow_off = static_cast<uint32_t>(row_in_tile) * 128u;
    const uint32_t xsw = static_cast<uint32_t>(row_in_tile & 7);
    uint32_t iter = 0, nchunk = 0;
    for (int tile = first_tile; tile < p.tiles_m; tile += tile_step, ++iter) {
      const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
      const int64_t row = static_cast<int64_t>(tile) * BM + row_in_tile;
      const bool row_ok = row < p.m;
      const uint32_t taddr = tmem_base + as * RN_BN + (static_cast<uint32_t>(q * 32) << 16) + chalf * 128;
      const int col_base = static_cast<int>(rank) * RN_BN + chalf * 128;
      float* red_stage = red + as * (2 * RN_MAXC * BM);
      // arm this tile's two exchange rounds (one arming thread per CTA)
      if (warp == EPI_WARP0 && lane == 0) {
        if (two_norms) mbar_arrive_expect_tx(&red_bar[as * 2 + 0], red_bytes);
        if (norm_out) mbar_arrive_expect_tx(&red_bar[as * 2 + 1], red_bytes);
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();

      auto exchange = [&](int round, float partial) -> float {
        // column half 1 hands its partial to half 0 through shared memory; half 0 pushes the CTA's sum of squares to
        // every CTA of the cluster (incl. itself); everybody then gathers the C partials in a fixed order
        if (chalf == 1) pairbuf[round * BM + row_in_tile] = partial;
        named_bar_sync(3, EPI_WARPS * 32);
        if (chalf == 0) {
          const float mine = partial + pairbuf[round * BM + row_in_tile];
          float* slot = red_stage + round * (RN_MAXC * BM) + static_cast<int>(rank) * BM + row_in_tile;
          const uint32_t slot_addr = smem_u32(slot), bar_addr = smem_u32(&red_bar[as * 2 + round]);
          for (int dst = 0; dst < C; ++dst) st_async_f32(mapa_u32(slot_addr, dst), mine, mapa_u32(bar_addr, dst));
        }
        mbar_wait(&red_bar[as * 2 + round], aphase);
        const float* base = red_stage + round * (RN_MAXC * BM) + row_in_tile;
        float total = 0.f;
        for (int src = 0; src < C; ++src) total += base[src * BM];  // same order in every CTA
        return total;
      };

      float rs_a = 1.0f;
      if (two_norms) {
        float ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[i]);
            ss = fmaf(v, v, ss);
          }
        }
        const float tot = exchange(0, row_ok ? ss : 0.f);
        rs_a = 1.0f / sqrtf(tot * inv_n + p.eps);
      }
      // pass 2: y = a * rs_a * w_post + x (or a + x) chunk by chunk through the staged x tiles; y replaces x in the
      // chunk buffer and leaves with a TMA store; y also replaces a in TMEM for pass 3; statistics of y
      float ssy = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c, ++nchunk) {
        const uint32_t slot = nchunk % RN_XSLOTS, xphase = (nchunk / RN_XSLOTS) & 1;
        const int bi = chalf * RN_XSLOTS + static_cast<int>(slot);
        uint8_t* buf = xbuf + bi * RN_XBUF_BYTES;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = col_base + c * 32;
        if (two_norms) {
          const float4* w4 = reinterpret_cast<const float4*>(wbuf + chalf * 128 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = w4[i];
            r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) * rs_a * w.x);
            r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) * rs_a * w.y);
            r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) * rs_a * w.z);
            r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) * rs_a * w.w);
          }
        }
        mbar_wait(&xfull_bar[bi], xphase);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4* xp = reinterpret_cast<float4*>(buf + xrow_off + ((static_cast<uint32_t>(i) ^ xsw) << 4));
          const float4 xv = *xp;
          const float y0 = __uint_as_float(r[4 * i + 0]) + xv.x, y1 = __uint_as_float(r[4 * i + 1]) + xv.y;
          const float y2 = __uint_as_float(r[4 * i + 2]) + xv.z, y3 = __uint_as_float(r[4 * i + 3]) + xv.w;
          *xp = make_float4(y0, y1, y2, y3);
          r[4 * i + 0] = __float_as_uint(y0), r[4 * i + 1] = __float_as_uint(y1);
          r[4 * i + 2] = __float_as_uint(y2), r[4 * i + 3] = __float_as_uint(y3);
          if (row_ok) ssy += y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3;
        }
        fence_proxy_async_smem();                     // generic-proxy writes of y -> visible to the TMA store
        named_bar_sync(1 + chalf, 4 * 32);            // the four warps of this column half finished the chunk
        if (store_issuer) {
          tma_store_2d(&p.tma_y, buf, col0, tile * BM);  // rows beyond m are clipped by the tensor map
          tma_store_commit();
          if (c > 0) {  // the previous chunk's buffer has been read by its store: hand it back to the x producer
            bulk_wait_read<1>();
            mbar_arrive(&xempty_bar[chalf * RN_XSLOTS + static_cast<int>((nchunk - 1) % RN_XSLOTS)]);
          }
          if (c == 3) {
            bulk_wait_read<0>();
            mbar_arrive(&xempty_bar[bi]);
          }
        }
        if (p.yn != nullptr) {
          if (norm_out) {
            tmem_st_32x32(taddr + c * 32, r);  // y replaces a in the accumulator columns for pass 3
          } else if (row_ok) {
            GemmParams gp = {};
            gp.d = p.yn, gp.ldd = (p.yn_dtype == TSFMX_DT_BF16_SPLIT ? 2 : 1) * static_cast<int64_t>(p.n);
            gp.d_dtype = p.yn_dtype, gp.n_store = p.n, gp.split_off = p.n, gp.vec_ok = 1, gp.split_k = 1;
            epilogue_chunk(gp, r, row, col0);
          }
        }
      }
      if (norm_out) {
        tmem_st_wait();
        const float tot = exchange(1, row_ok ? ssy : 0.f);
        const float rs_y = 1.0f / sqrtf(tot * inv_n + p.eps);
        GemmParams gp = {};
        gp.d = p.yn, gp.ldd = (p.yn_dtype == TSFMX_DT_BF16_SPLIT ? 2 : 1) * static_cast<int64_t>(p.n);
        gp.d_dtype = p.yn_dtype, gp.n_store = p.n, gp.split_off = p.n, gp.vec_ok = 1, gp.split_k = 1;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = col_base + c * 32;
          const float4* w4 = reinterpret_cast<const float4*>(wbuf + RN_BN + chalf * 128 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = w4[i];
            r[4 * i + 0] = __float_as_uint(w.x * (__uint_as_float(r[4 * i + 0]) * rs_y));
            r[4 * i + 1] = __float_as_uint(w.y * (__uint_as_float(r[4 * i + 1]) * rs_y));
            r[4 * i + 2] = __float_as_uint(w.z * (__uint_as_float(r[4 * i + 2]) * rs_y));
            r[4 * i + 3] = __float_as_uint(w.w * (__uint_as_float(r[4 * i + 3]) * rs_y));
          }
          if (row_ok) epilogue_chunk(gp, r, row, col0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
    if (store_issuer) tma_store_wait<0>();  // all y stores complete before the CTA may exit
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// bf16 [rows, cols] with row stride ld (elements); box = 64 x box_rows, 128B swizzle.
int make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return TSFMX_ERR_NO_DEVICE;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) base=%p rows=%lld cols=%lld ld=%lld", (int)r, base,
              (long long)rows, (long long)cols, (long long)ld);
    return TSFMX_ERR_CUDA;
  }
  return TSFMX_OK;
}

// fp32 [rows, cols] contiguous rows; box = 32 columns (128 B) x 128 rows, 128B swizzle: the x / y chunks of
// gemm_rownorm.  Rows beyond `rows` read as zero and are clipped on store.
int make_tmap_f32_chunk(CUtensorMap* map, const void* base, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return TSFMX_ERR_NO_DEVICE;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 4};
  cuuint32_t box[2] = {RN_XCHUNK, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32 chunk) failed (CUresult %d) base=%p rows=%lld cols=%lld", (int)r, base,
              (long long)rows, (long long)cols);
    return TSFMX_ERR_CUDA;
  }
  return TSFMX_OK;
}

template <int BN, int CG, bool TOKEN_MAJOR = false>
int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN, CG>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, CG, TOKEN_MAJOR>;
  static bool attr_set_dev[64] = {false};  // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.split_k;
  int units = num_sms() / CG;  // CTAs (or CTA pairs) resident at once: persistent grid
  if (units > total) units = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (CG == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return TSFMX_ERR_CUDA;
  }
  return check_last_launch("gemm_bf16_tcgen05");
}

int g_force_cta_group = 0;  // test hook: 0 = auto, 1 / 2 = force
int g_split_k_mode = 0;     // tune key 4: 0 = auto, 1 = never split K, n > 1 = force n splits where splitting is legal

// split-K for short-and-wide problems (weight gradients: M, N = features, K = tokens): a 1280 x 1280 output is 25
// tiles for 74 CTA pairs.  Only for a plain fp32 store (or an in-place accumulation into D), where the partial sums
// can be added with vector reductions into a zero-filled D, and only from K = 4096 up: the order of those
// reductions varies from run to run, which is fine for a gradient summed over tokens but would take away the
// run-to-run bit-reproducibility of the model GEMMs (K <= 3840 in every adapter here).
int choose_split_k(GemmParams& p, int cg, bool plain_f32_store, bool in_place, cudaStream_t stream) {
  p.split_k = 1;
  const int units = cg == 2 ? num_sms() / 2 : num_sms();
  const int base = p.tiles_m * p.tiles_n;
  int min_nkb = 1 << 30;
  for (int s = 0; s < p.num_xseg; ++s) min_nkb = p.xseg[s].nkb < min_nkb ? p.xseg[s].nkb : min_nkb;
  if (g_split_k_mode != 1 && base * 10 < units * 9 && plain_f32_store && (min_nkb >= 64 || g_split_k_mode > 1)) {
    double best = static_cast<double>(base) / units;  // no split: one partial wave
    for (int sk = 2; sk <= 16 && min_nkb / sk >= 16; ++sk) {  // at least 16 k-blocks (1024 of K) per split
      const int tiles = base * sk, waves = (tiles + units - 1) / units;
      const double kbs = static_cast<double>(min_nkb) / sk;
      const double score = static_cast<double>(tiles) / (waves * units) * kbs / (kbs + 4.0);
      if (score > best * 1.03) best = score, p.split_k = sk;
    }
    if (g_split_k_mode > 1) p.split_k = g_split_k_mode < min_nkb ? g_split_k_mode : min_nkb;
  }
  if (p.split_k > 1) {
    p.residual = nullptr;  // in-place accumulation: the reductions add to what D already holds
    if (!in_place) {
      const cudaError_t e = cudaMemset2DAsync(p.d, static_cast<size_t>(p.ldd) * 4, 0, static_cast<size_t>(p.n_store) * 4,
                                              static_cast<size_t>(p.m), stream);
      if (e != cudaSuccess) {
        set_error("gemm: cudaMemset2DAsync: %s", cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
    }
  }
  return TSFMX_OK;
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_gemm_set_cta_group(int cg) {
  if (cg < 0 || cg > 2) {
    set_error("cta_group must be 0 (auto), 1 or 2");
    return TSFMX_ERR_INVALID_ARGUMENT;
  }
  g_force_cta_group = cg;
  return TSFMX_OK;
}

extern "C" int tsfmx_gemm_set_split_k(int mode) {
  if (mode < 0 || mode > 16) {
    set_error("split_k mode must be 0 (auto), 1 (never) or 2..16 (forced where legal)");
    return TSFMX_ERR_INVALID_ARGUMENT;
  }
  g_split_k_mode = mode;
  return TSFMX_OK;
}

extern "C" int tsfmx_gemm(const tsfmx_gemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(a != nullptr, "gemm: args is NULL");
  TSFMX_REQUIRE(a->m > 0 && a->n > 0, "gemm: m (%lld) and n (%d) must be positive", (long long)a->m, a->n);
  TSFMX_REQUIRE(a->n % 8 == 0, "gemm: n (%d) must be a multiple of 8", a->n);
  TSFMX_REQUIRE(a->num_segments == 1 || a->num_segments == 2, "gemm: num_segments must be 1 or 2");
  TSFMX_REQUIRE(a->precision == TSFMX_PREC_BF16 || a->precision == TSFMX_PREC_BF16X3, "gemm: bad precision");
  TSFMX_REQUIRE(a->d != nullptr, "gemm: d is NULL");
  TSFMX_REQUIRE(a->d_dtype >= TSFMX_DT_F32 && a->d_dtype <= TSFMX_DT_BF16_SPLIT, "gemm: bad d_dtype");
  TSFMX_REQUIRE(a->act >= TSFMX_ACT_NONE && a->act <= TSFMX_ACT_RELU_GRAD, "gemm: bad act");
  TSFMX_REQUIRE(!(a->act == TSFMX_ACT_SILU_GRAD || a->act == TSFMX_ACT_RELU_GRAD) ||
                    (a->aux != nullptr && (a->aux_dtype == TSFMX_DT_F32 || a->aux_dtype == TSFMX_DT_BF16) &&
                     a->ld_aux >= (a->n_store > 0 ? a->n_store : a->n)),
                "gemm: *_GRAD activations need aux (f32 or bf16) with ld_aux >= n_store");
  TSFMX_REQUIRE(a->pre_act == nullptr || ((a->pre_act_dtype == TSFMX_DT_F32 || a->pre_act_dtype == TSFMX_DT_BF16) &&
                                          a->ld_pre >= (a->n_store > 0 ? a->n_store : a->n)),
                "gemm: pre_act must be f32 or bf16 with ld_pre >= n_store");
  TSFMX_REQUIRE(a->m < (int64_t(1) << 31) - 256, "gemm: m too large");
  const bool split = a->precision == TSFMX_PREC_BF16X3;

  int cg = g_force_cta_group;
  if (cg == 0) cg = 2;
  const int64_t rows_per_tile = BM * cg;
  const int bn = 256;

  GemmParams p = {};
  p.m = a->m;
  p.n = a->n;
  p.tiles_m = static_cast<int32_t>((a->m + rows_per_tile - 1) / rows_per_tile);
  p.tiles_n = (a->n + bn - 1) / bn;
  p.act = a->act;
  p.d_dtype = a->d_dtype;
  p.bias = a->bias;
  p.row_scale = a->row_scale;
  p.row_shift = a->row_shift;
  p.residual = a->residual;
  p.ldr = a->ldr;
  p.d = a->d;
  p.ldd = a->ldd;
  p.aux = a->aux;
  p.ld_aux = a->ld_aux;
  p.aux_dtype = a->aux_dtype;
  p.pre_act = a->pre_act;
  p.ld_pre = a->ld_pre;
  p.pre_dtype = a->pre_act_dtype;
  p.n_store = a->n_store > 0 ? a->n_store : a->n;
  TSFMX_REQUIRE(p.n_store <= a->n, "gemm: n_store (%d) > n (%d)", p.n_store, a->n);
  p.split_off = a->split_off > 0 ? a->split_off : a->n;
  TSFMX_REQUIRE(!(a->residual != nullptr) || a->ldr >= p.n_store, "gemm: ldr too small");
  TSFMX_REQUIRE(a->ldd >= p.n_store, "gemm: ldd too small");

  const int elem = a->d_dtype == TSFMX_DT_F32 ? 4 : 2;
  bool vec = (reinterpret_cast<uintptr_t>(a->d) % 16 == 0) && ((a->ldd * elem) % 16 == 0);
  if (a->d_dtype == TSFMX_DT_BF16_SPLIT) vec = vec && (p.split_off % 8 == 0);
  if (a->residual != nullptr) vec = vec && (reinterpret_cast<uintptr_t>(a->residual) % 16 == 0) && (a->ldr % 4 == 0);
  if (a->bias != nullptr) vec = vec && (reinterpret_cast<uintptr_t>(a->bias) % 16 == 0);
  p.vec_ok = vec ? 1 : 0;
  {
    const int ae = a->aux_dtype == TSFMX_DT_F32 ? 4 : 2, pe = a->pre_act_dtype == TSFMX_DT_F32 ? 4 : 2;
    p.aux_vec_ok = a->aux != nullptr && reinterpret_cast<uintptr_t>(a->aux) % 16 == 0 && (a->ld_aux * ae) % 16 == 0;
    p.pre_vec_ok = a->pre_act != nullptr && reinterpret_cast<uintptr_t>(a->pre_act) % 16 == 0 && (a->ld_pre * pe) % 16 == 0;
  }

  int nx = 0;
  for (int s = 0; s < a->num_segments; ++s) {
    const tsfmx_gemm_segment& sg = a->seg[s];
    TSFMX_REQUIRE(sg.a != nullptr && sg.b != nullptr, "gemm: segment %d has NULL operand", s);
    TSFMX_REQUIRE(sg.k > 0 && sg.k % BK == 0, "gemm: segment %d k (%d) must be a positive multiple of %d", s, sg.k, BK);
    TSFMX_REQUIRE(sg.lda % 8 == 0 && sg.ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
    TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(sg.a) % 16 == 0 && reinterpret_cast<uintptr_t>(sg.b) % 16 == 0,
                  "gemm: operands must be 16-byte aligned");
    const int64_t cols = split ? 2 * static_cast<int64_t>(sg.k) : sg.k;
    TSFMX_REQUIRE(sg.lda >= cols && sg.ldb >= cols, "gemm: lda/ldb smaller than the stored row length");
    int rc = make_tmap_bf16(&p.tma_a[s], sg.a, a->m, cols, sg.lda, BM);
    if (rc != TSFMX_OK) return rc;
    rc = make_tmap_bf16(&p.tma_b[s], sg.b, a->n, cols, sg.ldb, bn / cg);
    if (rc != TSFMX_OK) return rc;
    const int nkb = sg.k / BK;
    if (!split) {
      p.xseg[nx++] = XSeg{s, s, 0, 0, nkb};
    } else {
      p.xseg[nx++] = XSeg{s, s, sg.k, 0, nkb};  // lo * hi   (small terms first)
      p.xseg[nx++] = XSeg{s, s, 0, sg.k, nkb};  // hi * lo
      p.xseg[nx++] = XSeg{s, s, 0, 0, nkb};     // hi * hi
    }
  }
  p.num_xseg = nx;

  {
    const bool in_place = a->residual != nullptr && a->residual == a->d && a->ldr == a->ldd;
    const bool plain = a->d_dtype == TSFMX_DT_F32 && a->act == TSFMX_ACT_NONE && a->bias == nullptr &&
                       a->row_scale == nullptr && a->pre_act == nullptr && (a->residual == nullptr || in_place) &&
                       reinterpret_cast<uintptr_t>(a->d) % 16 == 0 && a->ldd % 4 == 0;
    const int rc = choose_split_k(p, cg, plain, in_place, stream);
    if (rc != TSFMX_OK) return rc;
  }

  if (cg == 2) return launch_gemm<256, 2>(p, stream);
  return launch_gemm<256, 1>(p, stream);
}

extern "C" int tsfmx_gemm_rownorm(const tsfmx_gemm_segment* seg, int64_t m, int32_t n, int32_t precision,
                                  const float* w_post, const float* w_next, const float* x, float* y,
                                  int32_t yn_dtype, void* yn, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(seg != nullptr && seg->a != nullptr && seg->b != nullptr, "gemm_rownorm: NULL operand");
  TSFMX_REQUIRE(m > 0 && m < (int64_t(1) << 31) - 256, "gemm_rownorm: bad m");
  TSFMX_REQUIRE(n > 0 && n % RN_BN == 0 && n / RN_BN >= 2 && n / RN_BN <= RN_MAXC,
                "gemm_rownorm: n (%d) must be 512, 768, 1024 or 1280", n);
  TSFMX_REQUIRE(precision == TSFMX_PREC_BF16 || precision == TSFMX_PREC_BF16X3, "gemm_rownorm: bad precision");
  TSFMX_REQUIRE(x != nullptr && y != nullptr, "gemm_rownorm: x and y are required");
  TSFMX_REQUIRE(yn == nullptr || (yn_dtype >= TSFMX_DT_F32 && yn_dtype <= TSFMX_DT_BF16_SPLIT), "gemm_rownorm: bad yn_dtype");
  TSFMX_REQUIRE(seg->k > 0 && seg->k % BK == 0, "gemm_rownorm: k (%d) must be a positive multiple of %d", seg->k, BK);
  TSFMX_REQUIRE(seg->lda % 8 == 0 && seg->ldb % 8 == 0, "gemm_rownorm: lda/ldb must be multiples of 8 elements");
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  TSFMX_REQUIRE(al16(seg->a) && al16(seg->b) && al16(x) && al16(y) && al16(yn) && al16(w_post) && al16(w_next),
                "gemm_rownorm: pointers must be 16-byte aligned");
  const bool split = precision == TSFMX_PREC_BF16X3;
  const int64_t cols = split ? 2 * static_cast<int64_t>(seg->k) : seg->k;
  TSFMX_REQUIRE(seg->lda >= cols && seg->ldb >= cols, "gemm_rownorm: lda/ldb smaller than the stored row length");

  RownormParams p = {};
  p.m = m;
  p.n = n;
  p.tiles_m = static_cast<int32_t>((m + BM - 1) / BM);
  p.cluster = n / RN_BN;
  p.a_issuers = p.cluster >= 4 ? 4 : 2;
  p.yn_dtype = yn_dtype;
  p.w_post = w_post;
  p.w_next = w_next;
  p.yn = yn;
  p.eps = eps;
  {
    int rcx = make_tmap_f32_chunk(&p.tma_x, x, m, n);
    if (rcx != TSFMX_OK) return rcx;
    rcx = make_tmap_f32_chunk(&p.tma_y, y, m, n);
    if (rcx != TSFMX_OK) return rcx;
  }
  int rc = make_tmap_bf16(&p.tma_a, seg->a, m, cols, seg->lda, BM / p.a_issuers);
  if (rc != TSFMX_OK) return rc;
  rc = make_tmap_bf16(&p.tma_b, seg->b, n, cols, seg->ldb, RN_BN);
  if (rc != TSFMX_OK) return rc;
  const int nkb = seg->k / BK;
  int nx = 0;
  if (!split) {
    p.xseg[nx++] = XSeg{0, 0, 0, 0, nkb};
  } else {
    p.xseg[nx++] = XSeg{0, 0, seg->k, 0, nkb};
    p.xseg[nx++] = XSeg{0, 0, 0, seg->k, nkb};
    p.xseg[nx++] = XSeg{0, 0, 0, 0, nkb};
  }
  p.num_xseg = nx;

  constexpr int SMEM = RN_STAGES * (BM * BK * 2 + RN_BN * BK * 2) + 2 * RN_XSLOTS * RN_XBUF_BYTES +
                       (2 * 2 * RN_MAXC * BM + 2 * BM + 2 * RN_BN) * 4 + 1024 + 256;
  static_assert(SMEM <= 227 * 1024, "gemm_rownorm: shared memory budget");
  auto kern = gemm_rownorm_tcgen05_kernel;
  static bool attr_set_dev[64] = {false};  // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
      set_error("gemm_rownorm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent: as many clusters as the device can hold at once
  static int max_clusters_dev[64][RN_MAXC + 1] = {{0}};
  int* max_clusters = max_clusters_dev[current_device()];
  if (max_clusters[p.cluster] == 0) {
    cfg.gridDim = dim3(p.cluster * 64);
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    if (e != cudaSuccess || nc <= 0) {
      (void)cudaGetLastError();
      nc = num_sms() / p.cluster;
    }
    max_clusters[p.cluster] = nc;
  }
  int clusters = max_clusters[p.cluster];
  if (clusters > p.tiles_m) clusters = p.tiles_m;
  cfg.gridDim = dim3(clusters * p.cluster);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    set_error("gemm_rownorm launch failed: %s", cudaGetErrorString(e));
    return TSFMX_ERR_CUDA;
  }
  return check_last_launch("gemm_rownorm_tcgen05");
}

extern "C" int tsfmx_gemm_wgrad(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t rows, int32_t n_out,
                                int32_t k_in, float* dw, int64_t ld_dw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(dy != nullptr && x != nullptr && dw != nullptr, "gemm_wgrad: NULL operand");
  TSFMX_REQUIRE(rows > 0 && rows < (int64_t(1) << 31) - 64, "gemm_wgrad: bad token count %lld", (long long)rows);
  TSFMX_REQUIRE(n_out > 0 && k_in > 0 && n_out % 8 == 0 && k_in % 8 == 0,
                "gemm_wgrad: n_out (%d) and k_in (%d) must be positive multiples of 8", n_out, k_in);
  TSFMX_REQUIRE(ld_dy >= n_out && ld_x >= k_in && ld_dy % 8 == 0 && ld_x % 8 == 0,
                "gemm_wgrad: row strides must cover the row and be multiples of 8 elements");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(dy) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(dw) % 16 == 0 && ld_dw % 4 == 0 && ld_dw >= k_in,
                "gemm_wgrad: operands must be 16-byte aligned (ld_dw a multiple of 4, >= k_in)");
  GemmParams p = {};
  p.m = n_out;
  p.n = k_in;
  p.tiles_m = (n_out + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (k_in + 255) / 256;
  p.act = TSFMX_ACT_NONE;
  p.d_dtype = TSFMX_DT_F32;
  p.d = dw;
  p.ldd = ld_dw;
  p.n_store = k_in;
  p.split_off = k_in;
  p.vec_ok = 1;
  // boxes of 64 features x 64 tokens straight out of the [tokens, features] matrices
  int rc = make_tmap_bf16(&p.tma_a[0], dy, rows, n_out, ld_dy, 64);
  if (rc != TSFMX_OK) return rc;
  rc = make_tmap_bf16(&p.tma_b[0], x, rows, k_in, ld_x, 64);
  if (rc != TSFMX_OK) return rc;
  p.xseg[0] = XSeg{0, 0, 0, 0, static_cast<int32_t>((rows + BK - 1) / BK)};
  p.num_xseg = 1;
  rc = choose_split_k(p, 2, true, false, stream);
  if (rc != TSFMX_OK) return rc;
  return launch_gemm<256, 2, true>(p, stream);
}

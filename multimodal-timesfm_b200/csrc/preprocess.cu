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
__global__ void __launch_bounds__(WARPS * 32) timesfm_patchify_norm_generic_kernel(
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
// Running-statistics helpers of the TimesFM kernels
// ----------------------------------------------------------------------------------------
struct RunStats {
  float n, mu, sigma;
};

// ----------------------------------------------------------------------------------------
// Warp-private TMA kernel (context <= 4096; longer or unaligned rows take the generic kernel at the top).  A
// block-wide version of it left 240 of its 256 threads at a barrier while one thread per series ran the dependent merge
// chain (ncu: 5.9 barrier-stall cycles per issued instruction, 24 % of the warp slots occupied; profiles/
// r1e_ncu_patchify_blockwide.md).  Here every WARP owns its tiles — about 32 patches: 2 series
// at context 512, one series from context 1024 — with its own two-deep ring of bulk copies and its own mbarriers, and
// no block-wide barrier after start-up: while a few lanes of one warp walk their merge chain, the scheduler runs the
// statistics and the token stores of the other warps.  Slots are series-major (slot of patch p of the tile at p).
// ----------------------------------------------------------------------------------------
constexpr int TFW_WARPS = 8;

// a / n, correctly rounded, for a divisor whose correctly rounded reciprocal r is known (n is a patch count, a small
// positive integer): two FMA refinement steps of q = a * r (Markstein); operands outside the range where the exact
// remainder argument holds take the IEEE division.
__device__ __forceinline__ float div_known_recip(float a, float n, float r) {
  const float aa = fabsf(a);
  if (aa < 0x1p100f && (aa > 0x1p-100f || a == 0.f)) {
    const float q0 = a * r;
    const float q1 = fmaf(fmaf(-n, q0, a), r, q0);
    return fmaf(fmaf(-n, q1, a), r, q1);
  }
  return __fdiv_rn(a, n);
}

// One step of the reference's update_running_stats - (n, mu, sigma) of the union of the running set and one patch -
// with the new count and its reciprocal supplied (the counts do not depend on the values, so they are scanned and
// inverted by all lanes in parallel before the dependent chain starts)
__device__ __forceinline__ RunStats merge_stats_r(RunStats run, float inc_n, float inc_mu, float inc_sigma, float new_n,
                                                  float r) {
  const float new_n_safe = new_n == 0.f ? 1.f : new_n;
  float new_mu = div_known_recip(__fadd_rn(__fmul_rn(run.n, run.mu), __fmul_rn(inc_mu, inc_n)), new_n_safe, r);
  if (new_n == 0.f) new_mu = 0.f;
  const float d1 = __fsub_rn(run.mu, new_mu), d2 = __fsub_rn(inc_mu, new_mu);
  const float t1 = __fmul_rn(run.n, __fmul_rn(run.sigma, run.sigma));
  const float t2 = __fmul_rn(inc_n, __fmul_rn(inc_sigma, inc_sigma));
  const float t3 = __fmul_rn(run.n, __fmul_rn(d1, d1));
  const float t4 = __fmul_rn(inc_n, __fmul_rn(d2, d2));
  float new_var = div_known_recip(__fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4), new_n_safe, r);
  if (new_n == 0.f) new_var = 0.f;
  return RunStats{new_n, new_mu, sqrtf(fmaxf(new_var, 0.f))};
}

// Statistics of the union of two disjoint sets (the same update, with the count and its IEEE reciprocal formed here):
// the operator of the parallel scan.
__device__ __forceinline__ RunStats merge_sets(RunStats a, RunStats b) {
  const float new_n = __fadd_rn(a.n, b.n);  // exact: small integers
  return merge_stats_r(a, b.n, b.mu, b.sigma, new_n, __frcp_rn(new_n == 0.f ? 1.f : new_n));
}
// the same for the scan steps, where either side may be an empty set (nothing to round then)
__device__ __forceinline__ RunStats merge_sets_scan(RunStats a, RunStats b) {
  if (a.n == 0.f) return b;
  if (b.n == 0.f) return a;
  return merge_sets(a, b);
}
__device__ __forceinline__ RunStats shfl_up_stats(RunStats v, int d) {
  return RunStats{__shfl_up_sync(0xffffffffu, v.n, d), __shfl_up_sync(0xffffffffu, v.mu, d),
                  __shfl_up_sync(0xffffffffu, v.sigma, d)};
}

template <int OUT>
__global__ void __launch_bounds__(TFW_WARPS * 32) timesfm_patchify_norm_warp_kernel(
    const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t batch, int context, int warps,
    int warp_bytes, int g, int stages, void* tokens, float* __restrict__ mu_out, float* __restrict__ sigma_out,
    uint8_t* __restrict__ patch_mask_out, int32_t* __restrict__ num_masked_out) {
  // g = series per tile (a power of two <= 32 with N <= 4 * 32 / g), stages = 1 or 2 tiles in flight
  extern __shared__ __align__(128) uint8_t smem_tfw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pg = lane >> 3, q = lane & 7;
  const int N = context >> 5;
  const int stage_bytes = (g * context * 5 + 127) & ~127;
  uint8_t* wbase = smem_tfw + warp * warp_bytes;
  float4* slots = reinterpret_cast<float4*>(wbase + stages * stage_bytes);  // {count, mean, sigma, padded flag} per patch
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slots + g * N);
  const int64_t num_tiles = (batch + g - 1) / g;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * warps + warp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * warps;

  auto issue = [&](int64_t tile, int stage) {
    const int64_t b0 = tile * g;
    const uint32_t cnt = static_cast<uint32_t>(batch - b0 < g ? batch - b0 : g);
    uint8_t* dst = wbase + stage * stage_bytes;
    mbar_arrive_expect_tx(&full_bar[stage], cnt * context * 5u);
    bulk_load_1d(dst, x + b0 * context, cnt * context * 4u, &full_bar[stage]);
    bulk_load_1d(dst + g * context * 4, mask + b0 * context, cnt * context, &full_bar[stage]);
  };
  if (lane == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
    if (first < num_tiles) issue(first, 0);
    if (stages > 1 && first + stride < num_tiles) issue(first + stride, 1);
  }
  __syncwarp();

  const int rot = lane + (lane >> 3);  // chunk rotation of phase A (bank groups), as in the block-wide kernel
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tile = first; tile < num_tiles; tile += stride) {
    const int64_t b0 = tile * g;
    const int cnt = static_cast<int>(batch - b0 < g ? batch - b0 : g);
    const int P = cnt * N;
    const float* sx = reinterpret_cast<const float*>(wbase + stage * stage_bytes);
    const uint8_t* sm = wbase + stage * stage_bytes + g * context * 4;
    mbar_wait(&full_bar[stage], parity);

    // ---- phase A: statistics of every patch, one lane per patch
    for (int p = lane; p < P; p += 32) {
      const float4* xp = reinterpret_cast<const float4*>(sx + p * 32);
      const uint32_t* mp = reinterpret_cast<const uint32_t*>(sm + p * 32);
      float xv[32];
      uint32_t mw[8];
      uint32_t any_padded = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = (j + rot) & 7;
        const float4 v = xp[ch];
        mw[j] = mp[ch];
        any_padded |= mw[j];
        xv[4 * j + 0] = v.x, xv[4 * j + 1] = v.y, xv[4 * j + 2] = v.z, xv[4 * j + 3] = v.w;
      }
      float c, inc_mu, sq = 0.f;
      if (any_padded == 0) {
        // fully observed patch (every patch of the reference's own callers, which pass all-False masks, and all but
        // the leading patches of a left-padded series): no per-element mask decoding - 3 instead of ~10 instructions
        // per element; same summation order as the general path
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) sum += xv[e];
        c = 32.f;
        inc_mu = sum * 0.03125f;  // == sum / 32 exactly
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = xv[e] - inc_mu;
          sq = fmaf(d, d, sq);
        }
      } else {
        uint32_t mbits = 0;  // bit e = element e of the patch is padded
        float sum = 0.f;
        c = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool padded = ((mw[j] >> (8 * k)) & 0xffu) != 0;
            mbits |= (padded ? 1u : 0u) << (4 * j + k);
            c += padded ? 0.f : 1.f;
            sum += padded ? 0.f : xv[4 * j + k];
          }
        }
        inc_mu = c == 0.f ? 0.f : __fdiv_rn(sum, c == 0.f ? 1.f : c);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = ((mbits >> e) & 1u) ? 0.f : xv[e] - inc_mu;
          sq = fmaf(d, d, sq);
        }
      }
      const float c_safe = c == 0.f ? 1.f : c;
      // the patch counts as padded iff its LAST element is padded (timesfm.py:97)
      const float flag = sm[p * 32 + 31] ? 1.f : 0.f;
      slots[p] = make_float4(c, inc_mu, c == 0.f ? 0.f : sqrtf(fmaxf(__fdiv_rn(sq, c_safe), 0.f)), flag);
    }
    __syncwarp();
    // ---- phase B: running statistics; the slot becomes {cumulative mu, cumulative sigma, 1 / safe sigma, padded flag}
    {
      // parallel scan.  The 32 / g lanes of a series own E <= 4 consecutive patches each: fold them locally, scan the
      // lane aggregates with shuffles (log2 steps of one set-union each), then apply the exclusive prefix - 2 E +
      // log2(32 / g) dependent merges with all lanes busy.  The folds this replaces kept one lane per series busy for N
      // dependent merges (ctx <= 512) or ran 24 dependent merges per series with 8 or 1 lanes active (blocked fold, ctx
      // 2048), plus a count pre-pass: ~1 600 of the kernel's 3 450 warp instructions per series at ctx 2048
      // (profiles/r2ac_ncu_patchify_ctx2048.md).  Every statistic is still the union of exact sub-set statistics
      // (update_running_stats' own formula); only the association order of the fp32 roundings differs from the
      // reference's patch-by-patch loop (timesfm.py:63-66) - a few 1e-7, inside the 2e-6 bound the tests hold.
      const int L = 32 / g, E = (N + L - 1) / L;
      const int sr = lane / L, ls = lane - sr * L, k0 = ls * E;
      const bool act = sr < cnt;
      RunStats loc[4];
      float flag[4];
      RunStats st = {0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        flag[e] = 0.f;
        if (e < E && k0 + e < N && act) {
          const float4 inc = slots[sr * N + k0 + e];
          flag[e] = inc.w;
          st = merge_sets(st, RunStats{inc.x, inc.y, inc.z});
        }
        loc[e] = st;
      }
      RunStats agg = st;
      for (int d = 1; d < L; d <<= 1) {
        const RunStats up = shfl_up_stats(agg, d);
        if (ls >= d) agg = merge_sets_scan(up, agg);
      }
      RunStats pre = shfl_up_stats(agg, 1);
      if (ls == 0) pre = RunStats{0.f, 0.f, 0.f};
      int masked = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (e < E && k0 + e < N && act) {
          const RunStats out = merge_sets_scan(pre, loc[e]);
          slots[sr * N + k0 + e] = make_float4(out.mu, out.sigma, 0.f, flag[e]);
          masked += flag[e] != 0.f ? 1 : 0;
        }
      }
      for (int d = 1; d < L; d <<= 1) masked += __shfl_xor_sync(0xffffffffu, masked, d);
      if (num_masked_out != nullptr && act && ls == 0) num_masked_out[b0 + sr] = masked;
    }
    __syncwarp();
    // ---- mu / sigma / patch mask out, coalesced (the tile's [cnt, N] block is contiguous in [B, N]); the
    //      reciprocal of the safe sigma is taken here, one lane per patch
    for (int i = lane; i < P; i += 32) {
      float4 st = slots[i];
      st.z = __fdiv_rn(1.0f, st.y < 1e-6f ? 1.f : st.y);
      slots[i] = st;
      if (mu_out != nullptr) mu_out[b0 * N + i] = st.x;
      if (sigma_out != nullptr) sigma_out[b0 * N + i] = st.y;
      if (patch_mask_out != nullptr) patch_mask_out[b0 * N + i] = st.w != 0.f ? 1 : 0;
    }
    __syncwarp();
    // ---- phase C: RevIN with the cumulative stats of the own patch, zero the padded points, emit tokens
    //      (8 lanes per patch: every store instruction writes whole 128-byte lines)
    for (int p = pg; p < P; p += 4) {
      const float4 st = slots[p];
      const float4 v = *reinterpret_cast<const float4*>(sx + p * 32 + 4 * q);
      const uint32_t mk = *reinterpret_cast<const uint32_t*>(sm + p * 32 + 4 * q);
      const float xv[4] = {v.x, v.y, v.z, v.w};
      float val[4], msk[4];
      if (mk == 0) {  // four observed points: no mask decoding
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) msk[kk] = 0.f, val[kk] = (xv[kk] - st.x) * st.z;
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const bool padded = ((mk >> (8 * kk)) & 0xffu) != 0;
          msk[kk] = padded ? 1.f : 0.f;
          val[kk] = padded ? 0.f : (xv[kk] - st.x) * st.z;
        }
      }
      store_token_quad<OUT>(tokens, b0 * N + p, q, val, msk);
    }
    __syncwarp();  // every lane is done with this stage and with the slots
    if (lane == 0) {
      const int64_t next = tile + static_cast<int64_t>(stages) * stride;
      if (next < num_tiles) issue(next, stage);
    }
    if (++stage == stages) { stage = 0; parity ^= 1; }
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

// asinh(x) = sign(x) log(|x| + sqrt(x^2 + 1)); odd Taylor series below 1/8 where the log form cancels.  The log form
// runs on the SFU (sqrt.approx + lg2.approx: ~1.6e-7 absolute error on a result >= 0.125, i.e. <= 1.3e-6 relative)
// and both forms are evaluated branch-free: ~14 instructions per element instead of ~45 for logf + IEEE sqrtf.
__device__ __forceinline__ float fast_asinh(float x) {
  const float t = fabsf(x);
  const float t2 = t * t;
  const float small = t * fmaf(t2, fmaf(t2, fmaf(t2, -0.044642857f, 0.075f), -0.16666667f), 1.0f);
  float root, lg;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(root) : "f"(t2 + 1.0f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t + root));
  const float big = lg * 0.6931471805599453f;
  return copysignf(t < 0.125f ? small : big, x);
}

// Fast path (context % 16 == 0, no left padding): one warp per series, lane l owns float4 #(l + 32 j) — four
// consecutive lanes own one 16-step patch — the row stays in registers across the three passes (mean, variance,
// emit) and every global access is a 16-byte streaming load / store.
template <int NV, int OUT>  // NV float4 per lane cached (context <= 128 * NV)
__global__ void __launch_bounds__(WARPS * 32) chronos2_patchify_norm_fast_kernel(
    const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t batch, int context, int use_arcsinh,
    float time_scale, int out_cols, void* out, uint8_t* __restrict__ attn_mask, float* __restrict__ loc_out,
    float* __restrict__ scale_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = context >> 2;
  const int num_patches = context >> 4;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    const uint8_t* mr = mask + b * context;
    float4 v[NV];
    uint32_t mk[NV];
    float sum = 0.f, cnt = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int f = lane + 32 * j;
      if (f < nvec) {
        v[j] = ld_stream_f4(xr + 4 * f);
        mk[j] = ld_stream_u32(mr + 4 * f);
      } else {
        v[j] = make_float4(NAN, NAN, NAN, NAN);
        mk[j] = 0x01010101u;
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!isnan(xv[k])) { sum += xv[k]; cnt += 1.f; }
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    const float loc = cnt > 0.f ? __fdiv_rn(sum, cnt) : 0.f;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!isnan(xv[k])) { const float d = xv[k] - loc; sq += d * d; }
    }
    sq = warp_sum(sq);
    float scale = cnt > 0.f ? sqrtf(__fdiv_rn(sq, cnt)) : 1.f;
    if (scale == 0.f) scale = 1e-5f;
    if (lane == 0) {
      if (loc_out != nullptr) loc_out[b] = loc;
      if (scale_out != nullptr) scale_out[b] = scale;
    }
    const int row_mul = OUT == TSFMX_DT_BF16_SPLIT ? 2 * out_cols : out_cols;
    const float inv_scale = __fdiv_rn(1.0f, scale);
    const float inv_time_scale = __fdiv_rn(1.0f, time_scale);
    // a power-of-two time scale (8192 for Chronos-2) makes the reciprocal multiply exact
    const bool tenc_exact = (__float_as_uint(time_scale) & 0x007fffffu) == 0u;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int f = lane + 32 * j;
      const bool live = f < nvec;
      const int patch = f >> 2, qd = f & 3;
      const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      float tenc[4], val[4], msk[4];
      int any = 0;
      const float e0 = static_cast<float>(4 * f - context);  // integers below 2^24: e0 + k is exact
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tenc[k] = tenc_exact ? (e0 + static_cast<float>(k)) * inv_time_scale
                             : __fdiv_rn(e0 + static_cast<float>(k), time_scale);
        const bool obs = ((mk[j] >> (8 * k)) & 0xffu) == 0;
        msk[k] = obs ? 1.f : 0.f;
        any |= obs ? 1 : 0;
        float sv = (xv[k] - loc) * inv_scale;
        if (use_arcsinh) sv = fast_asinh(sv);
        val[k] = obs ? sv : 0.f;
      }
      // attention mask of the patch = any observed point among its 16 steps (4 consecutive lanes)
      any |= __shfl_xor_sync(0xffffffffu, any, 1);
      any |= __shfl_xor_sync(0xffffffffu, any, 2);
      if (live) {
        if (attn_mask != nullptr && qd == 0) attn_mask[b * num_patches + patch] = static_cast<uint8_t>(any);
        const int64_t row = (b * num_patches + patch) * row_mul;
        if constexpr (OUT == TSFMX_DT_F32) {
          float* o = reinterpret_cast<float*>(out) + row;
          st_stream_f4(o + 4 * qd, make_float4(tenc[0], tenc[1], tenc[2], tenc[3]));
          st_stream_f4(o + 16 + 4 * qd, make_float4(val[0], val[1], val[2], val[3]));
          st_stream_f4(o + 32 + 4 * qd, make_float4(msk[0], msk[1], msk[2], msk[3]));
          if (out_cols >= 64) st_stream_f4(o + 48 + 4 * qd, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row;
          if constexpr (OUT == TSFMX_DT_BF16) {
            st_stream_u2(o + 4 * qd, make_uint2(pack_bf16x2(tenc[0], tenc[1]), pack_bf16x2(tenc[2], tenc[3])));
            st_stream_u2(o + 16 + 4 * qd, make_uint2(pack_bf16x2(val[0], val[1]), pack_bf16x2(val[2], val[3])));
            st_stream_u2(o + 32 + 4 * qd, make_uint2(pack_bf16x2(msk[0], msk[1]), pack_bf16x2(msk[2], msk[3])));
            if (out_cols >= 64) st_stream_u2(o + 48 + 4 * qd, make_uint2(0u, 0u));
          } else {
            uint2 h, l;
            split_bf16x2(tenc[0], tenc[1], h.x, l.x);
            split_bf16x2(tenc[2], tenc[3], h.y, l.y);
            st_stream_u2(o + 4 * qd, h);
            st_stream_u2(o + out_cols + 4 * qd, l);
            split_bf16x2(val[0], val[1], h.x, l.x);
            split_bf16x2(val[2], val[3], h.y, l.y);
            st_stream_u2(o + 16 + 4 * qd, h);
            st_stream_u2(o + out_cols + 16 + 4 * qd, l);
            st_stream_u2(o + 32 + 4 * qd, make_uint2(pack_bf16x2(msk[0], msk[1]), pack_bf16x2(msk[2], msk[3])));
            st_stream_u2(o + out_cols + 32 + 4 * qd, make_uint2(0u, 0u));
            if (out_cols >= 64) {
              st_stream_u2(o + 48 + 4 * qd, make_uint2(0u, 0u));
              st_stream_u2(o + out_cols + 48 + 4 * qd, make_uint2(0u, 0u));
            }
          }
        }
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

template <int NE>  // elements cached per lane (context <= 32 * NE); NE == 0: re-read the row
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

  // lane l owns elements l, l + 32, ...: 128-byte coalesced loads, 256-byte coalesced int64 id stores
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    int64_t* idr = ids + b * (context + 1);
    uint8_t* amr = attn_mask + b * (context + 1);
    float cache[NE > 0 ? NE : 1];
    // sum(|x|) is accumulated in fp64 and rounded to fp32 once, so the scale (and hence every id)
    // does not depend on the reduction order.
    double sum = 0.0;
    float cnt = 0.f;
    if (NE > 0) {
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int e = lane + 32 * k;
        cache[k] = e < context ? ld_stream_f32(xr + e) : NAN;
      }
#pragma unroll
      for (int k = 0; k < NE; ++k)
        if (!isnan(cache[k])) { sum += static_cast<double>(fabsf(cache[k])); cnt += 1.f; }
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
    if (NE > 0) {
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int e = lane + 32 * k;
        if (e < context) {
          idr[e] = tok(cache[k]);
          amr[e] = isnan(cache[k]) ? 0 : 1;
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

// Search of the boundary table kept in shared memory between two sentinels (-inf below, NaN above), so that it
// needs no bounds checks: the uniform-grid guess is verified against its two neighbouring entries and only a miss walks
// the table (exact for any ascending table).
__device__ __forceinline__ int bucketize_right_sentinel(const float* __restrict__ sp, int nb, float v, float g_mul,
                                                        float g_add) {
  // sp[0] = -inf, sp[1 + i] = boundaries[i], sp[nb + 1] = NaN; returns #boundaries <= v (v is not NaN)
  float g = fmaf(v, g_mul, g_add);
  g = fminf(fmaxf(g, 0.f), static_cast<float>(nb));
  int i = static_cast<int>(g);  // candidate count in [0, nb]
  const float lo = sp[i], hi = sp[i + 1];
  if (!(lo <= v) || hi <= v) {
    while (!(sp[i] <= v)) --i;
    while (sp[i + 1] <= v) ++i;
  }
  return i;
}

// Staged variant (the default for context % 4 == 0, context <= 2048).  Same arithmetic as the vectorised kernel
// above, but the tokens and mask bytes of a series go through a per-warp shared-memory row first, so that what
// leaves the SM is re-tiled to the ROW's alignment in global memory: every 16-byte id store is one aligned pair
// (an int64 row of context + 1 entries starts on an 8-byte boundary for every other series) and every mask store
// is one aligned 32-bit word (a mask row of context + 1 bytes starts at any byte).  A warp store instruction then
// covers 512 / 128 contiguous bytes instead of half-filled sectors at a 32-byte stride.  The per-element
// magnitude test of the hoisted-reciprocal division is gone: the quotient is only trusted when it lands strictly
// inside a cell of the uniform grid (frac in [1/256, 255/256]), and every other element takes the IEEE division
// and the exact table search, so a quotient whose refinement under/overflowed can never pick the token.
constexpr int T5_CHUNK = 512;  // elements staged per warp at a time (4 float4 per lane)

template <bool MULTI>  // MULTI: context > T5_CHUNK, the row is read twice (the second time from L2)
__global__ void __launch_bounds__(WARPS * 32) chronos_t5_tokenize_staged_kernel(
    const float* __restrict__ x, int64_t batch, int context, const float* __restrict__ boundaries, int nb,
    int n_special, int n_tokens, int pad_id, int eos_id, int64_t* __restrict__ ids,
    uint8_t* __restrict__ attn_mask, float* __restrict__ scale_out) {
  extern __shared__ __align__(16) float sp[];  // [nb + 2] table between sentinels, then the per-warp staging rows
  __shared__ int s_nonuniform;
  if (threadIdx.x == 0) s_nonuniform = 0;
  for (int i = threadIdx.x; i < nb + 2; i += blockDim.x)
    sp[i] = i == 0 ? -INFINITY : (i == nb + 1 ? NAN : boundaries[i - 1]);
  __syncthreads();
  const float step = (sp[nb - 1] - sp[2]) / static_cast<float>(nb - 3);
  const float inv_step = static_cast<float>(nb - 3) / (sp[nb - 1] - sp[2]);
  const float g_mul = inv_step, g_add = fmaf(-sp[2], inv_step, 2.0f);
  {
    int bad = !(step > 0.f) || nb > 8190;
    for (int i = 1 + threadIdx.x; i <= nb - 2; i += blockDim.x) {
      const float ideal = fmaf(static_cast<float>(i - 1), step, sp[2]);
      bad |= !(fabsf(sp[1 + i] - ideal) <= step * (1.0f / 2048.0f));
    }
    if (bad) s_nonuniform = 1;
  }
  __syncthreads();
  const bool uniform = s_nonuniform == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = context >> 2;
  const int row = context + 1;
  const int t_max = n_tokens - 1;
  const int nchunks = MULTI ? (context + T5_CHUNK - 1) / T5_CHUNK : 1;
  // staging row of this warp, reused chunk after chunk: token of chunk-relative element e at st_tok[4 + e]
  // (e = -1: last token of the previous chunk), mask byte of chunk-relative position p at byte 4 + p of st_msk32
  // (word 0: last four bytes of the previous chunk)
  constexpr int TOK_WORDS = T5_CHUNK + 8, MSK_WORDS = (T5_CHUNK + 12) >> 2, WARP_WORDS = (TOK_WORDS + MSK_WORDS + 3) & ~3;
  uint32_t* st_tok = reinterpret_cast<uint32_t*>(sp + ((nb + 2 + 3) & ~3)) + warp * WARP_WORDS;
  uint32_t* st_msk32 = st_tok + TOK_WORDS;

  const uint64_t keep = MULTI ? l2_policy_evict_last() : 0, drop = MULTI ? l2_policy_evict_first() : 0;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * WARPS + warp; b < batch;
       b += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + b * context;
    int64_t* idr = ids + b * row;
    uint8_t* amr = attn_mask + b * row;
    float4 v[4];
    // sum(|x|) accumulated in fp64 and rounded to fp32 once: the scale (hence every id) is order independent
    double sum = 0.0;
    float cnt = 0.f;
    // MULTI reads the row twice: the first read asks L2 to keep the lines (evict_last), the second one releases
    // them (evict_first); without the hints the 2:1 stream of stores pushed the rows out in between and ncu showed
    // 1.03 GB of DRAM reads for 0.54 GB of input
    auto load_chunk = [&](float4 (&dst)[4], int c, uint64_t policy) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = c * (T5_CHUNK / 4) + lane + 32 * j;
        dst[j] = f >= nvec ? make_float4(NAN, NAN, NAN, NAN)
                           : (MULTI ? ld_hint_f4(xr + 4 * f, policy) : ld_stream_f4(xr + 4 * f));
      }
    };
    auto accumulate = [&](const float4 (&src)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xv[4] = {src[j].x, src[j].y, src[j].z, src[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool nan = isnan(xv[k]);
          sum += nan ? 0.0 : static_cast<double>(fabsf(xv[k]));
          cnt += nan ? 0.f : 1.f;
        }
      }
    };
    if (MULTI) {  // two chunks (eight 16-byte loads per lane) in flight
      float4 w[4];
      for (int c = 0; c < nchunks; c += 2) {
        load_chunk(v, c, keep);
        load_chunk(w, c + 1, keep);  // all-NaN beyond the row
        accumulate(v);
        accumulate(w);
      }
      load_chunk(v, 0, drop);  // second read, first chunk: in flight during the reduction below
    } else {
      load_chunk(v, 0, drop);
      accumulate(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt = warp_sum(cnt);
    float scale = __fdiv_rn(static_cast<float>(sum), cnt);  // 0/0 -> NaN -> 1 below
    if (!(scale > 0.f)) scale = 1.f;
    const float rscale = __frcp_rn(scale);
    // the refined quotient is only meaningful for a scale whose reciprocal is a normal number
    const float g_hi = (uniform && scale > 0x1p-60f && scale < 0x1p60f) ? static_cast<float>(nb - 1) : -1.f;
    if (lane == 0 && scale_out != nullptr) scale_out[b] = scale;
    const int s = static_cast<int>((reinterpret_cast<uintptr_t>(idr) >> 3) & 1);  // row starts in the upper half of a pair
    const int a = static_cast<int>(reinterpret_cast<uintptr_t>(amr) & 3);
    const int sh = 8 * (4 - a);

    for (int c = 0; c < nchunks; ++c) {
      const int len = min(T5_CHUNK, context - c * T5_CHUNK);  // elements of this chunk (a multiple of 4)
      const bool last = c == nchunks - 1;
      if (last && lane == 0) {
        st_tok[4 + len] = static_cast<uint32_t>(eos_id);
        st_msk32[1 + (len >> 2)] = 1u;  // eos byte
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = lane + 32 * j;  // float4 index inside the chunk
        if (4 * f < len) {
          const float xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          uint32_t t[4];
          uint32_t mbytes = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool nan = isnan(xv[k]);
            const float av = nan ? 0.f : xv[k];
            // av / scale through the hoisted correctly rounded reciprocal and two FMA refinement steps (Markstein)
            const float q0 = av * rscale;
            const float q1 = fmaf(fmaf(-scale, q0, av), rscale, q0);
            const float q = fmaf(fmaf(-scale, q1, av), rscale, q1);
            const float g = fmaf(q, g_mul, g_add);
            const float fl = floorf(g);
            const float frac = g - fl;
            int tk;
            if (g >= 2.0f && g < g_hi && frac >= (1.0f / 256.0f) && frac <= (255.0f / 256.0f))
              tk = static_cast<int>(fl);
            else
              tk = bucketize_right_sentinel(sp, nb, __fdiv_rn(av, scale), g_mul, g_add);
            tk = max(0, min(t_max, tk + n_special));
            t[k] = nan ? static_cast<uint32_t>(pad_id) : static_cast<uint32_t>(tk);
            mbytes |= (nan ? 0u : 1u) << (8 * k);
          }
          *reinterpret_cast<uint4*>(st_tok + 4 + 4 * f) = make_uint4(t[0], t[1], t[2], t[3]);
          st_msk32[1 + f] = mbytes;
        }
      }
      if (MULTI && !last) load_chunk(v, c + 1, drop);  // second read of the row (an L2 hit), in flight during the stores
      __syncwarp();
      const int top = last ? len : len - 1;  // highest chunk-relative index that holds a value (eos included)
      // ---- ids: aligned 16-byte pairs (token e0 = 2k - s and its successor)
      {
        uint32_t* base = reinterpret_cast<uint32_t*>(idr + c * T5_CHUNK - s);  // 16-byte aligned
        const int pairs = ((top + s) >> 1) + 1;
        for (int k = lane; k < pairs; k += 32) {
          const int e0 = 2 * k - s;
          const uint32_t t0 = st_tok[4 + e0], t1 = st_tok[5 + e0];
          uint32_t* o = base + 4 * k;
          const bool lo_ok = e0 >= 0 || c > 0, hi_ok = e0 + 1 <= top;
          if (lo_ok && hi_ok) {
            if (MULTI) st_hint_u4(o, make_uint4(t0, 0u, t1, 0u), drop);  // the ids must not push the rows out of L2
            else *reinterpret_cast<uint4*>(o) = make_uint4(t0, 0u, t1, 0u);
          } else if (!lo_ok) {
            if (hi_ok) *reinterpret_cast<uint2*>(o + 2) = make_uint2(t1, 0u);
          } else if (last) {  // the row ends in the lower half of its pair
            *reinterpret_cast<uint2*>(o) = make_uint2(t0, 0u);
          }  // else: the pair straddles the chunk boundary and leaves with the next chunk
        }
      }
      // ---- mask: aligned 32-bit words; the first and last word of a row belong partly to its neighbours
      {
        uint8_t* base = amr + c * T5_CHUNK - a;
        const int words = ((top + a) >> 2) + 1;
        for (int k = lane; k < words; k += 32) {
          const int p0 = 4 * k - a;  // chunk-relative positions p0 .. p0 + 3
          const uint32_t w = __funnelshift_rc(st_msk32[k], st_msk32[k + 1], sh);
          const bool lo_ok = p0 >= 0 || c > 0, hi_ok = p0 + 3 <= top;
          if (lo_ok && hi_ok) {
            *reinterpret_cast<uint32_t*>(base + 4 * k) = w;
          } else if (hi_ok || last) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if ((p0 + i >= 0 || c > 0) && p0 + i <= top) base[4 * k + i] = static_cast<uint8_t>(w >> (8 * i));
          }  // else: the word straddles the chunk boundary and leaves with the next chunk
        }
      }
      __syncwarp();
      if (!last && lane == 0) {  // carry the chunk's tail into the slots in front of the next chunk
        st_tok[3] = st_tok[4 + T5_CHUNK - 1];
        st_msk32[0] = st_msk32[T5_CHUNK >> 2];
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

// tuning hooks (0 = default): series per warp tile / warps per block of the TimesFM kernel; g_t5_variant = 1 and
// g_tf_variant = 1 force the generic fallback kernels (A/B runs, tests of the fallbacks)
int g_tf_group = 0, g_tf_warps = 0, g_t5_variant = 0, g_tf_variant = 0;

template <int OUT>
int launch_timesfm_staged(const float* x, const uint8_t* mask, int64_t batch, int context, void* tokens, float* mu,
                          float* sigma, uint8_t* patch_mask, int32_t* num_masked, cudaStream_t stream) {
  const int N = context >> 5;
  if (context <= 4096) {  // warp-private tiles, no block-wide barriers
    // series per tile: phase B keeps one lane per series busy, so more series per tile means fewer (mostly idle)
    // warp instructions per series; the tile still has to leave room for >= 12 warps per SM
    // series per tile: a power of two (32 / g lanes scan one series, at most 4 patches per lane); the tile still has
    // to leave room for >= 12 warps per SM
    int want = g_tf_group > 0 ? g_tf_group : 2048 / context;
    int g = 1;
    while (2 * g <= want && 2 * g <= 32 && N <= 4 * (32 / (2 * g))) g *= 2;
    int stages = (g_tf_warps >> 8) > 0 ? (g_tf_warps >> 8) : (g * context * 5 <= 6 * 1024 ? 2 : 1);
    if (stages > 2) stages = 2;
    const int stage_bytes = (g * context * 5 + 127) & ~127;
    const int warp_bytes = (stages * stage_bytes + g * N * 16 + 16 + 127) & ~127;
    int warps = (g_tf_warps & 0xff) > 0 ? (g_tf_warps & 0xff) : TFW_WARPS;
    if (warps > TFW_WARPS) warps = TFW_WARPS;
    while (warps > 1 && warps * warp_bytes > 110 * 1024) --warps;  // at least two blocks per SM
    const int smem = warps * warp_bytes;
    auto kern = timesfm_patchify_norm_warp_kernel<OUT>;
    static int smem_set_dev[64] = {0};  // cudaFuncSetAttribute is per device
    int& smem_set = smem_set_dev[current_device()];
    if (smem > 48 * 1024 && smem > smem_set) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) {
        set_error("timesfm_patchify_norm: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        return TSFMX_ERR_CUDA;
      }
      smem_set = smem;
    }
    const int64_t tiles = (batch + g - 1) / g;
    const int64_t blocks = (tiles + warps - 1) / warps;
    int per_sm = (224 * 1024) / (smem + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm * warps > 48) per_sm = 48 / warps > 0 ? 48 / warps : 1;
    const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
    const int grid = static_cast<int>(blocks < cap ? blocks : cap);
    kern<<<grid, warps * 32, smem, stream>>>(x, mask, batch, context, warps, warp_bytes, g, stages, tokens, mu, sigma,
                                             patch_mask, num_masked);
    return check_last_launch("timesfm_patchify_norm");
  }
  set_error("timesfm_patchify_norm: context %d is beyond the staged kernel", context);
  return TSFMX_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_tune(int32_t key, int32_t value) {
  switch (key) {
    case 0: g_tf_group = value; return TSFMX_OK;
    case 1: g_tf_warps = value; return TSFMX_OK;
    case 2: g_t5_variant = value; return TSFMX_OK;
    case 3: g_tf_variant = value; return TSFMX_OK;
    case 5: g_t5_general_attention = value; return TSFMX_OK;
    default:
      set_error("tune: unknown key %d", key);
      return TSFMX_ERR_INVALID_ARGUMENT;
  }
}

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
  TSFMX_REQUIRE(tokens_dtype >= TSFMX_DT_F32 && tokens_dtype <= TSFMX_DT_BF16_SPLIT,
                "timesfm_patchify_norm: bad tokens_dtype %d", tokens_dtype);
  if (batch == 0) return TSFMX_OK;
  if (g_tf_variant == 0 && context <= 4096 && reinterpret_cast<uintptr_t>(mask) % 16 == 0) {
    switch (tokens_dtype) {
      case TSFMX_DT_F32:
        return launch_timesfm_staged<TSFMX_DT_F32>(x, mask, batch, context, tokens, mu, sigma, patch_mask, num_masked, stream);
      case TSFMX_DT_BF16:
        return launch_timesfm_staged<TSFMX_DT_BF16>(x, mask, batch, context, tokens, mu, sigma, patch_mask, num_masked, stream);
      default:
        return launch_timesfm_staged<TSFMX_DT_BF16_SPLIT>(x, mask, batch, context, tokens, mu, sigma, patch_mask, num_masked, stream);
    }
  }
  const int grid = grid_for_series(batch);
  const dim3 block(WARPS * 32);
  switch (tokens_dtype) {
    case TSFMX_DT_F32:
      timesfm_patchify_norm_generic_kernel<TSFMX_DT_F32><<<grid, block, 0, stream>>>(x, mask, batch, context, tokens, mu,
                                                                                    sigma, patch_mask, num_masked);
      break;
    case TSFMX_DT_BF16:
      timesfm_patchify_norm_generic_kernel<TSFMX_DT_BF16><<<grid, block, 0, stream>>>(x, mask, batch, context, tokens, mu,
                                                                                     sigma, patch_mask, num_masked);
      break;
    default:
      timesfm_patchify_norm_generic_kernel<TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(
          x, mask, batch, context, tokens, mu, sigma, patch_mask, num_masked);
      break;
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
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "chronos2_patchify_norm: bad out_dtype %d",
                out_dtype);
  const bool fast = patch == 16 && context % 16 == 0 && context <= 2048 && (out_cols == 48 || out_cols == 64) &&
                    reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(mask) % 4 == 0 &&
                    reinterpret_cast<uintptr_t>(patched) % 16 == 0;
  if (fast) {
#define TSFMX_C2_LAUNCH(NVV, DT)                                                                              \
  chronos2_patchify_norm_fast_kernel<NVV, DT><<<grid, block, 0, stream>>>(x, mask, batch, context, use_arcsinh, \
                                                                          time_encoding_scale, out_cols, patched, \
                                                                          attn_mask, loc, scale)
    if (context <= 512) {
      if (out_dtype == TSFMX_DT_F32) TSFMX_C2_LAUNCH(4, TSFMX_DT_F32);
      else if (out_dtype == TSFMX_DT_BF16) TSFMX_C2_LAUNCH(4, TSFMX_DT_BF16);
      else TSFMX_C2_LAUNCH(4, TSFMX_DT_BF16_SPLIT);
    } else if (context <= 1024) {
      if (out_dtype == TSFMX_DT_F32) TSFMX_C2_LAUNCH(8, TSFMX_DT_F32);
      else if (out_dtype == TSFMX_DT_BF16) TSFMX_C2_LAUNCH(8, TSFMX_DT_BF16);
      else TSFMX_C2_LAUNCH(8, TSFMX_DT_BF16_SPLIT);
    } else {
      if (out_dtype == TSFMX_DT_F32) TSFMX_C2_LAUNCH(16, TSFMX_DT_F32);
      else if (out_dtype == TSFMX_DT_BF16) TSFMX_C2_LAUNCH(16, TSFMX_DT_BF16);
      else TSFMX_C2_LAUNCH(16, TSFMX_DT_BF16_SPLIT);
    }
#undef TSFMX_C2_LAUNCH
    return check_last_launch("chronos2_patchify_norm");
  }
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
  // default: the staged kernel, any context that is a multiple of 4; everything else (odd contexts, unaligned rows,
  // tiny tables) takes the scalar kernel that re-reads the row
  const bool staged_ok = g_t5_variant == 0 && context % 4 == 0 && n_boundaries >= 4 && pad_id >= 0 &&
                         reinterpret_cast<uintptr_t>(x) % 16 == 0;
  if (staged_ok) {
    const size_t warp_words = ((T5_CHUNK + 8) + ((T5_CHUNK + 12) >> 2) + 3) & ~3;
    const size_t smem_staged = (static_cast<size_t>((n_boundaries + 2 + 3) & ~3) + WARPS * warp_words) * sizeof(float);
    static bool attr_done[64] = {};  // per device: an 8192-entry table plus the staging rows exceed 48 KB
    const int dev = current_device();
    if (smem_staged > 48 * 1024 && !attr_done[dev]) {
      if (cudaFuncSetAttribute(chronos_t5_tokenize_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               64 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(chronos_t5_tokenize_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               64 * 1024) != cudaSuccess) {
        set_error("chronos_t5_tokenize: cannot raise the dynamic shared memory limit");
        return TSFMX_ERR_CUDA;
      }
      attr_done[dev] = true;
    }
    if (context <= T5_CHUNK)
      chronos_t5_tokenize_staged_kernel<false><<<grid, block, smem_staged, stream>>>(
          x, batch, context, boundaries, n_boundaries, n_special, n_tokens, pad_id, eos_id, ids, attn_mask, scale);
    else
      chronos_t5_tokenize_staged_kernel<true><<<grid, block, smem_staged, stream>>>(
          x, batch, context, boundaries, n_boundaries, n_special, n_tokens, pad_id, eos_id, ids, attn_mask, scale);
    return check_last_launch("chronos_t5_tokenize");
  }
  chronos_t5_tokenize_kernel<0><<<grid, block, smem, stream>>>(x, batch, context, boundaries, n_boundaries, n_special,
                                                               n_tokens, pad_id, eos_id, ids, attn_mask, scale);
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

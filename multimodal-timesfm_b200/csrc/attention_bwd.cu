// Tensor-core backward of the TimesFM 2.5 attention core (fusion fine-tune step, reference tsfmx/trainer.py:200-219:
// the activation gradient has to cross every frozen decoder layer to reach the fusion weights).
//
// Same decomposition as the forward kernel (attention.cu): one warp owns one (series, head), N <= 64 patches.
//   stage      raw q / k / v and dO rows -> padded bf16 tiles (16-byte cp.async)
//   condition  q' = w_q * RMSNorm(RoPE(q)) * scale, k' = w_k * RMSNorm(RoPE(k)) in fp32, parked as bf16 next to the
//              raw tiles (the raw rows are needed again for the RMSNorm / RoPE backward)
//   pass A     per 16-query tile: S = Q'K'^T and dP = dO V^T on mma.sync.m16n8k16, softmax / delta / dS in
//              registers, dQ' = dS K' on the tensor cores; row statistics (max, 1/sum, delta) kept for pass B
//   pass B     per 16-key tile: the (query tile, key tile) blocks of P and dS are recomputed from the row
//              statistics, transposed in registers (movmatrix) and fed back as A operands: dV = P^T dO, dK' = dS^T Q'
//   un-condition  dq, dk through the per-dim scale, the RMSNorm and the inverse rotation (fp32, two lanes per row)
// Masking: causal + left-padded patches; a query row without any admissible key attends uniformly to all N keys and
// passes gradient to them, like the reference's additive finfo.min mask does under autograd.
#include <math.h>

#include "common.cuh"

namespace tsfmx {
namespace {

constexpr int BW_HD = 80;
constexpr int BW_LD = 88;   // bf16 tile row (176 B): ldmatrix rows land in distinct bank groups
constexpr int BW_LDF = 84;  // fp32 staging row (336 B)

__device__ __forceinline__ void bw_cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bw_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void bw_ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void bw_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// transpose of one 8x8 b16 block held in the mma fragment layout (lane = 4 * row + column pair)
__device__ __forceinline__ uint32_t bw_movmatrix(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float2 bw_pair(uint32_t w) {  // the two bf16 of a word as fp32
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// dst[0..20) += sum over the 16 lanes of equal parity of v[0..20) (lane bit 0 selects the destination and is not summed)
__device__ __forceinline__ void bw_fold20(const float (&v)[20], int lane, float* dst) {
  const bool up1 = (lane & 2) != 0, up2 = (lane & 4) != 0;
  float w10[10], w5[5];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float send = up1 ? v[i] : v[i + 10], keep = up1 ? v[i + 10] : v[i];
    w10[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float send = up2 ? w10[i] : w10[i + 5], keep = up2 ? w10[i + 5] : w10[i];
    w5[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    w5[i] += __shfl_xor_sync(0xffffffffu, w5[i], 8);
    w5[i] += __shfl_xor_sync(0xffffffffu, w5[i], 16);
  }
  if ((lane & 24) == 0) {
    float* d = dst + (up1 ? 10 : 0) + (up2 ? 5 : 0);
#pragma unroll
    for (int i = 0; i < 5; ++i) d[i] += w5[i];
  }
}

template <int NT>  // 16-row tiles: N <= 16 * NT
__global__ void __launch_bounds__(256) timesfm_attention_bwd_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, int64_t batch, int num_patches,
    int num_heads, const uint8_t* __restrict__ patch_mask, const int32_t* __restrict__ num_masked,
    const float* __restrict__ inv_freq, const float* __restrict__ q_ln_w, const float* __restrict__ k_ln_w,
    const float* __restrict__ q_scale, float eps, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dparams) {
  constexpr int ROWS = 16 * NT;
  constexpr int HALF = BW_HD / 2;
  constexpr int TILE = ROWS * BW_LD;    // bf16 elements
  constexpr int TILEF = ROWS * BW_LDF;  // floats
  constexpr int PER_WARP = 6 * TILE * 2 + 2 * TILEF * 4 + 4 * ROWS * 4;
  extern __shared__ __align__(16) uint8_t smem_bw[];
  float* s_rope = reinterpret_cast<float*>(smem_bw);  // [2 * ROWS][cos 40 | sin 40 | -sin 40], positions -N .. N-1
  float* s_wq = s_rope + 2 * ROWS * 3 * HALF;         // [80] q_ln_w * q_scale
  float* s_wk = s_wq + BW_HD;
  float* s_dw = s_wk + BW_HD;  // [warps][2 * 80] per-warp sums of dq' * qhat and dk' * khat (full fine-tune only)
  uint8_t* warp_base = reinterpret_cast<uint8_t*>(s_dw + 8 * 2 * BW_HD);
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = num_patches;
  __nv_bfloat16* sQraw = reinterpret_cast<__nv_bfloat16*>(warp_base + warp * PER_WARP);
  __nv_bfloat16* sKraw = sQraw + TILE;
  __nv_bfloat16* sQ = sKraw + TILE;  // conditioned q'
  __nv_bfloat16* sK = sQ + TILE;
  __nv_bfloat16* sV = sK + TILE;     // v, later the dv staging
  __nv_bfloat16* sDO = sV + TILE;
  float* sDQ = reinterpret_cast<float*>(sDO + TILE);  // dL/dq' fp32
  float* sDK = sDQ + TILEF;
  float* sMx = sDK + TILEF;  // per query row: max, 1 / sum, delta, dead flag
  float* sInv = sMx + ROWS;
  float* sDelta = sInv + ROWS;
  float* sDead = sDelta + ROWS;

  for (int i = threadIdx.x; i < 2 * N * HALF; i += blockDim.x) {
    const int p = i / HALF, f = i - p * HALF;
    float sn, cs;
    sincosf(static_cast<float>(p - N) * __ldg(inv_freq + f), &sn, &cs);
    float* trow = s_rope + p * 3 * HALF;
    trow[f] = cs, trow[HALF + f] = sn, trow[2 * HALF + f] = -sn;
  }
  for (int i = threadIdx.x; i < BW_HD; i += blockDim.x) {
    s_wq[i] = __ldg(q_ln_w + i) * __ldg(q_scale + i);
    s_wk[i] = __ldg(k_ln_w + i);
  }
  for (int i = threadIdx.x; i < 8 * 2 * BW_HD; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  float* my_dw = s_dw + warp * 2 * BW_HD;

  const int width = num_heads * BW_HD;
  const int64_t qkv_ld = 3 * static_cast<int64_t>(width);
  const int64_t total = batch * num_heads;
  const int g = lane >> 2, t = lane & 3;
  const int ntk = (N + 15) >> 4;
  const uint32_t sq_addr = smem_u32(sQ), sk_addr = smem_u32(sK), sv_addr = smem_u32(sV), sdo_addr = smem_u32(sDO);

  for (int64_t unit = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; unit < total;
       unit += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t b = unit / num_heads;
    const int h = static_cast<int>(unit - b * num_heads);
    const int nm = num_masked != nullptr ? num_masked[b] : 0;

    // ---- stage raw q / k / v and dO rows (10 x 16 B each); zero the rows beyond N
    const __nv_bfloat16* gbase = qkv + b * N * qkv_ld + h * BW_HD;
    const __nv_bfloat16* dobase = dout + b * N * static_cast<int64_t>(width) + h * BW_HD;
    for (int c = lane; c < N * 10; c += 32) {
      const int row = c / 10, ch = c - row * 10;
      bw_cp_async_16(sQraw + row * BW_LD + ch * 8, gbase + row * qkv_ld + ch * 8);
      bw_cp_async_16(sKraw + row * BW_LD + ch * 8, gbase + row * qkv_ld + width + ch * 8);
      bw_cp_async_16(sV + row * BW_LD + ch * 8, gbase + row * qkv_ld + 2 * width + ch * 8);
      bw_cp_async_16(sDO + row * BW_LD + ch * 8, dobase + static_cast<int64_t>(row) * width + ch * 8);
    }
    for (int c = lane; c < (ntk * 16 - N) * 10; c += 32) {
      const int row = N + c / 10, ch = c % 10;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(sQ + row * BW_LD + ch * 8) = z;
      *reinterpret_cast<uint4*>(sK + row * BW_LD + ch * 8) = z;
      *reinterpret_cast<uint4*>(sV + row * BW_LD + ch * 8) = z;
      *reinterpret_cast<uint4*>(sDO + row * BW_LD + ch * 8) = z;
    }
    uint64_t kmask = 0;  // bit j: key j exists and is not padded
    {
      const bool v0 = lane < N && (patch_mask == nullptr || patch_mask[b * N + lane] == 0);
      const bool v1 = lane + 32 < N && (patch_mask == nullptr || patch_mask[b * N + lane + 32] == 0);
      kmask = static_cast<uint64_t>(__ballot_sync(0xffffffffu, v0)) |
              (static_cast<uint64_t>(__ballot_sync(0xffffffffu, v1)) << 32);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    // ---- forward conditioning (lane pair (2r, 2r+1) owns row r, 20 rotary pairs per lane): raw -> q', k'
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
      const int r = rt * 16 + (lane >> 1);
      if (rt < ntk) {
        const int hf = lane & 1;
        const bool live = r < N;
        // packed fp32 pairs (dims 2j, 2j + 1 of this lane's 20): FFMA2 / FMUL2, as in the forward kernel
        const float* rope = s_rope + (live ? (r - nm + N) : 0) * 3 * HALF + 20 * hf;
        const __nv_bfloat16* qraw = sQraw + r * BW_LD + 20 * hf;
        const __nv_bfloat16* kraw = sKraw + r * BW_LD + 20 * hf;
        float2 q1[10], q2[10], k1[10], k2[10];
        float2 qss2 = make_float2(0.f, 0.f), kss2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const uint2 a1 = live ? *reinterpret_cast<const uint2*>(qraw + 4 * i) : make_uint2(0u, 0u);
          const uint2 a2 = live ? *reinterpret_cast<const uint2*>(qraw + HALF + 4 * i) : make_uint2(0u, 0u);
          const uint2 c1 = live ? *reinterpret_cast<const uint2*>(kraw + 4 * i) : make_uint2(0u, 0u);
          const uint2 c2 = live ? *reinterpret_cast<const uint2*>(kraw + HALF + 4 * i) : make_uint2(0u, 0u);
          const uint32_t aw1[2] = {a1.x, a1.y}, aw2[2] = {a2.x, a2.y}, cw1[2] = {c1.x, c1.y}, cw2[2] = {c2.x, c2.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * i + e;
            const float2 cs = *reinterpret_cast<const float2*>(rope + 2 * j);
            const float2 sn = *reinterpret_cast<const float2*>(rope + HALF + 2 * j);
            const float2 ns = *reinterpret_cast<const float2*>(rope + 2 * HALF + 2 * j);
            const float2 x1 = bw_pair(aw1[e]), x2 = bw_pair(aw2[e]), y1 = bw_pair(cw1[e]), y2 = bw_pair(cw2[e]);
            q1[j] = fma2(x2, ns, mul2(x1, cs));
            q2[j] = fma2(x1, sn, mul2(x2, cs));
            k1[j] = fma2(y2, ns, mul2(y1, cs));
            k2[j] = fma2(y1, sn, mul2(y2, cs));
            qss2 = fma2(q2[j], q2[j], fma2(q1[j], q1[j], qss2));
            kss2 = fma2(k2[j], k2[j], fma2(k1[j], k1[j], kss2));
          }
        }
        float qss = qss2.x + qss2.y, kss = kss2.x + kss2.y;
        qss += __shfl_xor_sync(0xffffffffu, qss, 1);
        kss += __shfl_xor_sync(0xffffffffu, kss, 1);
        const float qrs = rsqrtf(qss * (1.0f / BW_HD) + eps);
        const float krs = rsqrtf(kss * (1.0f / BW_HD) + eps);
        const float2 qrs2 = make_float2(qrs, qrs), krs2 = make_float2(krs, krs);
        const float* wq1 = s_wq + 20 * hf;
        const float* wk1 = s_wk + 20 * hf;
        __nv_bfloat16* qrow = sQ + r * BW_LD + 20 * hf;
        __nv_bfloat16* krow = sK + r * BW_LD + 20 * hf;
        if (live) {
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
            *reinterpret_cast<uint2*>(qrow + 4 * i) = make_uint2(o1[0], o1[1]);
            *reinterpret_cast<uint2*>(qrow + HALF + 4 * i) = make_uint2(o2[0], o2[1]);
            *reinterpret_cast<uint2*>(krow + 4 * i) = make_uint2(p1[0], p1[1]);
            *reinterpret_cast<uint2*>(krow + HALF + 4 * i) = make_uint2(p2[0], p2[1]);
          }
        }
      }
    }
    __syncwarp();

    // ---- pass A: per query tile -> row statistics, dQ'
#pragma unroll
    for (int qi = 0; qi < NT; ++qi) {
      if (qi < ntk) {
        uint32_t qa[5][4], da[5][4];
        {
          const int row = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int col = 8 * (lane >> 4);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {
            bw_ldmatrix_x4(sq_addr + (row * BW_LD + col + 16 * ks) * 2, qa[ks]);
            bw_ldmatrix_x4(sdo_addr + (row * BW_LD + col + 16 * ks) * 2, da[ks]);
          }
        }
        const int row0 = qi * 16 + g, row1 = row0 + 8;
        const bool dead0 = (kmask & ((2ull << row0) - 1ull)) == 0ull;
        const bool dead1 = (kmask & ((2ull << row1) - 1ull)) == 0ull;
        const bool any_dead = __any_sync(0xffffffffu, (dead0 && row0 < N) || (dead1 && row1 < N));
        const int kj_end = any_dead ? ntk : qi + 1;

        float s[NT][2][4], dp[NT][2][4];
#pragma unroll
        for (int kj = 0; kj < NT; ++kj) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[kj][nt][e] = 0.f, dp[kj][nt][e] = 0.f;
          if (kj < kj_end) {
            const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
            const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
              uint32_t kb[4], vb[4];
              bw_ldmatrix_x4(sk_addr + (key * BW_LD + col + 16 * ks) * 2, kb);
              bw_mma(s[kj][0], qa[ks], kb[0], kb[1]);
              bw_mma(s[kj][1], qa[ks], kb[2], kb[3]);
              bw_ldmatrix_x4(sv_addr + (key * BW_LD + col + 16 * ks) * 2, vb);
              bw_mma(dp[kj][0], da[ks], vb[0], vb[1]);
              bw_mma(dp[kj][1], da[ks], vb[2], vb[3]);
            }
          }
        }
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
        const float ml0 = mx0 * LOG2E_F, ml1 = mx1 * LOG2E_F;  // row maxima in the exp2 domain (kept for pass B)
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
        // delta = sum_j P_ij dP_ij ; dS = P (dP - delta)
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
        if (t == 0) {
          sMx[row0] = ml0, sInv[row0] = inv0, sDelta[row0] = dl0, sDead[row0] = dead0 ? 1.f : 0.f;
          sMx[row1] = ml1, sInv[row1] = inv1, sDelta[row1] = dl1, sDead[row1] = dead1 ? 1.f : 0.f;
        }
        float dq[10][4];
#pragma unroll
        for (int dt = 0; dt < 10; ++dt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dq[dt][e] = 0.f;
#pragma unroll
        for (int kj = 0; kj < NT; ++kj) {
          if (kj < kj_end) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[kj][0][0] * (dp[kj][0][0] - dl0), s[kj][0][1] * (dp[kj][0][1] - dl0));
            pa[1] = pack_bf16x2(s[kj][0][2] * (dp[kj][0][2] - dl1), s[kj][0][3] * (dp[kj][0][3] - dl1));
            pa[2] = pack_bf16x2(s[kj][1][0] * (dp[kj][1][0] - dl0), s[kj][1][1] * (dp[kj][1][1] - dl0));
            pa[3] = pack_bf16x2(s[kj][1][2] * (dp[kj][1][2] - dl1), s[kj][1][3] * (dp[kj][1][3] - dl1));
            const int key = kj * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int col = 8 * (lane >> 4);
#pragma unroll
            for (int dpp = 0; dpp < 5; ++dpp) {
              uint32_t kb[4];
              bw_ldmatrix_x4_trans(sk_addr + (key * BW_LD + col + 16 * dpp) * 2, kb);
              bw_mma(dq[2 * dpp], pa, kb[0], kb[1]);
              bw_mma(dq[2 * dpp + 1], pa, kb[2], kb[3]);
            }
          }
        }
#pragma unroll
        for (int dt = 0; dt < 10; ++dt) {
          *reinterpret_cast<float2*>(sDQ + row0 * BW_LDF + dt * 8 + 2 * t) = make_float2(dq[dt][0], dq[dt][1]);
          *reinterpret_cast<float2*>(sDQ + row1 * BW_LDF + dt * 8 + 2 * t) = make_float2(dq[dt][2], dq[dt][3]);
        }
      }
    }
    __syncwarp();

    // ---- pass B: per key tile -> dV, dK'
#pragma unroll
    for (int kj = 0; kj < NT; ++kj) {
      if (kj < ntk) {
        float dv[10][4], dk[10][4];
#pragma unroll
        for (int dt = 0; dt < 10; ++dt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dv[dt][e] = 0.f, dk[dt][e] = 0.f;
        uint32_t kf[5][4], vf[5][4];  // B operands of S = Q'K'^T and dP = dO V^T for this key tile
        {
          const int key = kj * 16 + (lane & 7) + 8 * (lane >> 4);
          const int col = 8 * ((lane >> 3) & 1);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {
            bw_ldmatrix_x4(sk_addr + (key * BW_LD + col + 16 * ks) * 2, kf[ks]);
            bw_ldmatrix_x4(sv_addr + (key * BW_LD + col + 16 * ks) * 2, vf[ks]);
          }
        }
#pragma unroll
        for (int qi = 0; qi < NT; ++qi) {
          if (qi < ntk) {
            const int row0 = qi * 16 + g, row1 = row0 + 8;
            const float dead0 = sDead[row0], dead1 = sDead[row1];
            const bool tile_dead = __any_sync(0xffffffffu, (dead0 != 0.f && row0 < N) || (dead1 != 0.f && row1 < N));
            if (qi >= kj || tile_dead) {
              float s2[2][4], dp2[2][4];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) s2[nt][e] = 0.f, dp2[nt][e] = 0.f;
              {
                const int row = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int col = 8 * (lane >> 4);
#pragma unroll
                for (int ks = 0; ks < 5; ++ks) {
                  uint32_t qa[4], da[4];
                  bw_ldmatrix_x4(sq_addr + (row * BW_LD + col + 16 * ks) * 2, qa);
                  bw_ldmatrix_x4(sdo_addr + (row * BW_LD + col + 16 * ks) * 2, da);
                  bw_mma(s2[0], qa, kf[ks][0], kf[ks][1]);
                  bw_mma(s2[1], qa, kf[ks][2], kf[ks][3]);
                  bw_mma(dp2[0], da, vf[ks][0], vf[ks][1]);
                  bw_mma(dp2[1], da, vf[ks][2], vf[ks][3]);
                }
              }
              const float mx0 = sMx[row0], mx1 = sMx[row1], inv0 = sInv[row0], inv1 = sInv[row1];
              const float dl0 = sDelta[row0], dl1 = sDelta[row1];
              float p[2][4], ds[2][4];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int key = kj * 16 + nt * 8 + 2 * t + (e & 1);
                  const int row = (e & 2) ? row1 : row0;
                  const bool dead = ((e & 2) ? dead1 : dead0) != 0.f;
                  const bool ok = kj <= qi && key <= row && ((kmask >> key) & 1ull);
                  float pv;
                  if (dead) pv = key < N ? 1.f : 0.f;
                  else pv = ok ? exp_sub(s2[nt][e], (e & 2) ? mx1 : mx0) : 0.f;  // sMx holds mx * log2e
                  pv *= (e & 2) ? inv1 : inv0;
                  if (row >= N) pv = 0.f;
                  p[nt][e] = pv;
                  ds[nt][e] = pv * (dp2[nt][e] - ((e & 2) ? dl1 : dl0));
                }
              // A operands = transposed blocks: rows = keys of tile kj, columns = queries of tile qi
              uint32_t pT[4], dsT[4];
              pT[0] = bw_movmatrix(pack_bf16x2(p[0][0], p[0][1]));   // (keys 0-7,  queries 0-7)
              pT[1] = bw_movmatrix(pack_bf16x2(p[1][0], p[1][1]));   // (keys 8-15, queries 0-7)
              pT[2] = bw_movmatrix(pack_bf16x2(p[0][2], p[0][3]));   // (keys 0-7,  queries 8-15)
              pT[3] = bw_movmatrix(pack_bf16x2(p[1][2], p[1][3]));   // (keys 8-15, queries 8-15)
              dsT[0] = bw_movmatrix(pack_bf16x2(ds[0][0], ds[0][1]));
              dsT[1] = bw_movmatrix(pack_bf16x2(ds[1][0], ds[1][1]));
              dsT[2] = bw_movmatrix(pack_bf16x2(ds[0][2], ds[0][3]));
              dsT[3] = bw_movmatrix(pack_bf16x2(ds[1][2], ds[1][3]));
              const int qrow = qi * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
              const int col = 8 * (lane >> 4);
#pragma unroll
              for (int dpp = 0; dpp < 5; ++dpp) {
                uint32_t ob[4], qb[4];
                bw_ldmatrix_x4_trans(sdo_addr + (qrow * BW_LD + col + 16 * dpp) * 2, ob);
                bw_mma(dv[2 * dpp], pT, ob[0], ob[1]);
                bw_mma(dv[2 * dpp + 1], pT, ob[2], ob[3]);
                bw_ldmatrix_x4_trans(sq_addr + (qrow * BW_LD + col + 16 * dpp) * 2, qb);
                bw_mma(dk[2 * dpp], dsT, qb[0], qb[1]);
                bw_mma(dk[2 * dpp + 1], dsT, qb[2], qb[3]);
              }
            }
          }
        }
        // v rows of this key tile are dead now: park dv there (bf16) for the coalesced store; dk' to fp32 staging
        __syncwarp();
        const int key0 = kj * 16 + g, key1 = key0 + 8;
#pragma unroll
        for (int dt = 0; dt < 10; ++dt) {
          *reinterpret_cast<uint32_t*>(sV + key0 * BW_LD + dt * 8 + 2 * t) = pack_bf16x2(dv[dt][0], dv[dt][1]);
          *reinterpret_cast<uint32_t*>(sV + key1 * BW_LD + dt * 8 + 2 * t) = pack_bf16x2(dv[dt][2], dv[dt][3]);
          *reinterpret_cast<float2*>(sDK + key0 * BW_LDF + dt * 8 + 2 * t) = make_float2(dk[dt][0], dk[dt][1]);
          *reinterpret_cast<float2*>(sDK + key1 * BW_LDF + dt * 8 + 2 * t) = make_float2(dk[dt][2], dk[dt][3]);
        }
        __syncwarp();
      }
    }
    __syncwarp();

    // ---- back through per-dim scale * RMSNorm and the rotation (lane pair per row); dq, dk overwrite q', k' tiles
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
      const int r = rt * 16 + (lane >> 1);
      if (rt < ntk) {
        const int hf = lane & 1;
        const bool live = r < N;
        const float* rope = s_rope + (live ? (r - nm + N) : 0) * 3 * HALF + 20 * hf;
        const __nv_bfloat16* qraw = sQraw + r * BW_LD + 20 * hf;
        const __nv_bfloat16* kraw = sKraw + r * BW_LD + 20 * hf;
        const float* dqp = sDQ + r * BW_LDF + 20 * hf;
        const float* dkp = sDK + r * BW_LDF + 20 * hf;
        const float* wq1 = s_wq + 20 * hf;
        const float* wk1 = s_wk + 20 * hf;
        auto ld2 = [](const float* p_) { return *reinterpret_cast<const float2*>(p_); };
        float2 q1[10], q2[10], k1[10], k2[10];
        float2 qss2 = make_float2(0.f, 0.f), kss2 = make_float2(0.f, 0.f);
        float2 cq2 = make_float2(0.f, 0.f), ck2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const uint2 a1 = live ? *reinterpret_cast<const uint2*>(qraw + 4 * i) : make_uint2(0u, 0u);
          const uint2 a2 = live ? *reinterpret_cast<const uint2*>(qraw + HALF + 4 * i) : make_uint2(0u, 0u);
          const uint2 c1 = live ? *reinterpret_cast<const uint2*>(kraw + 4 * i) : make_uint2(0u, 0u);
          const uint2 c2 = live ? *reinterpret_cast<const uint2*>(kraw + HALF + 4 * i) : make_uint2(0u, 0u);
          const uint32_t aw1[2] = {a1.x, a1.y}, aw2[2] = {a2.x, a2.y}, cw1[2] = {c1.x, c1.y}, cw2[2] = {c2.x, c2.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * i + e;
            const float2 cs = ld2(rope + 2 * j), sn = ld2(rope + HALF + 2 * j), ns = ld2(rope + 2 * HALF + 2 * j);
            const float2 x1 = bw_pair(aw1[e]), x2 = bw_pair(aw2[e]), y1 = bw_pair(cw1[e]), y2 = bw_pair(cw2[e]);
            q1[j] = fma2(x2, ns, mul2(x1, cs));
            q2[j] = fma2(x1, sn, mul2(x2, cs));
            k1[j] = fma2(y2, ns, mul2(y1, cs));
            k2[j] = fma2(y1, sn, mul2(y2, cs));
            qss2 = fma2(q2[j], q2[j], fma2(q1[j], q1[j], qss2));
            kss2 = fma2(k2[j], k2[j], fma2(k1[j], k1[j], kss2));
            // t = w * dL/dq' ; c = <rope(q), t>
            cq2 = fma2(q1[j], mul2(ld2(wq1 + 2 * j), ld2(dqp + 2 * j)), cq2);
            cq2 = fma2(q2[j], mul2(ld2(wq1 + HALF + 2 * j), ld2(dqp + HALF + 2 * j)), cq2);
            ck2 = fma2(k1[j], mul2(ld2(wk1 + 2 * j), ld2(dkp + 2 * j)), ck2);
            ck2 = fma2(k2[j], mul2(ld2(wk1 + HALF + 2 * j), ld2(dkp + HALF + 2 * j)), ck2);
          }
        }
        float qss = qss2.x + qss2.y, kss = kss2.x + kss2.y, cq = cq2.x + cq2.y, ck = ck2.x + ck2.y;
        qss += __shfl_xor_sync(0xffffffffu, qss, 1);
        kss += __shfl_xor_sync(0xffffffffu, kss, 1);
        cq += __shfl_xor_sync(0xffffffffu, cq, 1);
        ck += __shfl_xor_sync(0xffffffffu, ck, 1);
        const float rq = rsqrtf(qss * (1.0f / BW_HD) + eps), rk = rsqrtf(kss * (1.0f / BW_HD) + eps);
        const float fq = rq * rq * rq * cq * (1.0f / BW_HD), fk = rk * rk * rk * ck * (1.0f / BW_HD);
        if (dparams != nullptr) {
          // parameter gradients: the 16 lanes with the same half (one per row of the tile) hold contributions to the
          // same 40 dims.  Summing every value over all 16 lanes cost 4 shuffles per value - 1 480 of this kernel's
          // 4 835 warp instructions per (series, head) in the full fine-tune step (profiles/r2aa_ncu_finetune_backward.md)
          // - so the first two exchange steps HALVE the set instead (a lane keeps 10, then 5 of its 20 sums and
          // hands the others over), and only the last two steps are plain butterflies: 25 shuffles per 20 values.
          float v[20];  // one group of 20 at a time: the register file is full here
#pragma unroll
          for (int j = 0; j < 10; ++j) v[2 * j] = dqp[2 * j] * q1[j].x * rq, v[2 * j + 1] = dqp[2 * j + 1] * q1[j].y * rq;
          bw_fold20(v, lane, my_dw + 20 * hf);
#pragma unroll
          for (int j = 0; j < 10; ++j)
            v[2 * j] = dqp[HALF + 2 * j] * q2[j].x * rq, v[2 * j + 1] = dqp[HALF + 2 * j + 1] * q2[j].y * rq;
          bw_fold20(v, lane, my_dw + 20 * hf + HALF);
#pragma unroll
          for (int j = 0; j < 10; ++j) v[2 * j] = dkp[2 * j] * k1[j].x * rk, v[2 * j + 1] = dkp[2 * j + 1] * k1[j].y * rk;
          bw_fold20(v, lane, my_dw + BW_HD + 20 * hf);
#pragma unroll
          for (int j = 0; j < 10; ++j)
            v[2 * j] = dkp[HALF + 2 * j] * k2[j].x * rk, v[2 * j + 1] = dkp[HALF + 2 * j + 1] * k2[j].y * rk;
          bw_fold20(v, lane, my_dw + BW_HD + 20 * hf + HALF);
        }
        __nv_bfloat16* qrow = sQ + r * BW_LD + 20 * hf;
        __nv_bfloat16* krow = sK + r * BW_LD + 20 * hf;
        if (live) {
          const float2 rq2 = make_float2(rq, rq), rk2 = make_float2(rk, rk);
          const float2 nfq2 = make_float2(-fq, -fq), nfk2 = make_float2(-fk, -fk);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            uint32_t gq1[2], gq2[2], gk1[2], gk2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = 2 * i + e;
              const float2 cs = ld2(rope + 2 * j), sn = ld2(rope + HALF + 2 * j), ns = ld2(rope + 2 * HALF + 2 * j);
              // d/d rope(q) = rq (w dq') - rope(q) fq, then the transposed rotation
              const float2 a = fma2(mul2(ld2(wq1 + 2 * j), ld2(dqp + 2 * j)), rq2, mul2(q1[j], nfq2));
              const float2 bq = fma2(mul2(ld2(wq1 + HALF + 2 * j), ld2(dqp + HALF + 2 * j)), rq2, mul2(q2[j], nfq2));
              const float2 c = fma2(mul2(ld2(wk1 + 2 * j), ld2(dkp + 2 * j)), rk2, mul2(k1[j], nfk2));
              const float2 d = fma2(mul2(ld2(wk1 + HALF + 2 * j), ld2(dkp + HALF + 2 * j)), rk2, mul2(k2[j], nfk2));
              const float2 g1 = fma2(bq, sn, mul2(a, cs)), g2 = fma2(a, ns, mul2(bq, cs));
              const float2 h1 = fma2(d, sn, mul2(c, cs)), h2 = fma2(c, ns, mul2(d, cs));
              gq1[e] = pack_bf16x2(g1.x, g1.y), gq2[e] = pack_bf16x2(g2.x, g2.y);
              gk1[e] = pack_bf16x2(h1.x, h1.y), gk2[e] = pack_bf16x2(h2.x, h2.y);
            }
            *reinterpret_cast<uint2*>(qrow + 4 * i) = make_uint2(gq1[0], gq1[1]);
            *reinterpret_cast<uint2*>(qrow + HALF + 4 * i) = make_uint2(gq2[0], gq2[1]);
            *reinterpret_cast<uint2*>(krow + 4 * i) = make_uint2(gk1[0], gk1[1]);
            *reinterpret_cast<uint2*>(krow + HALF + 4 * i) = make_uint2(gk2[0], gk2[1]);
          }
        }
      }
    }
    __syncwarp();
    // ---- coalesced 16-byte stores of dq | dk | dv
    __nv_bfloat16* obase = dqkv + b * N * qkv_ld + h * BW_HD;
    for (int c = lane; c < N * 10; c += 32) {
      const int row = c / 10, ch = c - row * 10;
      __nv_bfloat16* o = obase + row * qkv_ld + ch * 8;
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(sQ + row * BW_LD + ch * 8);
      *reinterpret_cast<uint4*>(o + width) = *reinterpret_cast<const uint4*>(sK + row * BW_LD + ch * 8);
      *reinterpret_cast<uint4*>(o + 2 * width) = *reinterpret_cast<const uint4*>(sV + row * BW_LD + ch * 8);
    }
    __syncwarp();
  }
  if (dparams != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * BW_HD; i += blockDim.x) {
      float t2 = 0.f;
      for (int w2 = 0; w2 < warps_per_block; ++w2) t2 += s_dw[w2 * 2 * BW_HD + i];
      atomicAdd(dparams + i, t2);
    }
  }
}

template <int NT>
int launch_bwd_mma(const void* qkv, const void* dout, int64_t batch, int N, int H, const uint8_t* pm, const int32_t* nm,
                   const float* inv_freq, const float* qw, const float* kw, const float* qs, float eps, void* dqkv,
                   float* dparams, cudaStream_t stream) {
  constexpr int ROWS = 16 * NT;
  constexpr int per_warp = 6 * ROWS * BW_LD * 2 + 2 * ROWS * BW_LDF * 4 + 4 * ROWS * 4;
  constexpr int fixed = 2 * ROWS * 3 * 40 * 4 + 2 * BW_HD * 4 + 8 * 2 * BW_HD * 4;
  int wpb = (220 * 1024 - fixed) / per_warp;
  if (wpb > 8) wpb = 8;
  if (wpb < 1) {
    set_error("timesfm_attention_bwd: %d patches do not fit the tensor-core kernel", N);
    return TSFMX_ERR_UNSUPPORTED;
  }
  const int smem = fixed + wpb * per_warp;
  auto kern = timesfm_attention_bwd_mma_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("timesfm_attention_bwd: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return TSFMX_ERR_CUDA;
    }
  }
  const int64_t total = batch * H;
  const int64_t blocks = (total + wpb - 1) / wpb;
  const int per_sm = (224 * 1024) / (smem + 1024) > 0 ? (224 * 1024) / (smem + 1024) : 1;
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  kern<<<grid, wpb * 32, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                          reinterpret_cast<const __nv_bfloat16*>(dout), batch, N, H, pm, nm, inv_freq, qw,
                                          kw, qs, eps, reinterpret_cast<__nv_bfloat16*>(dqkv), dparams);
  return check_last_launch("timesfm_attention_bwd_mma");
}

}  // namespace

// Used by tsfmx_timesfm_attention_bwd (backward.cu) when qkv, dO and dqkv are all bf16 and N <= 64.
int launch_timesfm_attention_bwd_mma(const void* qkv, const void* dout, int64_t batch, int N, int H, const uint8_t* pm,
                                     const int32_t* nm, const float* inv_freq, const float* qw, const float* kw,
                                     const float* qs, float eps, void* dqkv, float* dparams, cudaStream_t stream) {
  if (N <= 16) return launch_bwd_mma<1>(qkv, dout, batch, N, H, pm, nm, inv_freq, qw, kw, qs, eps, dqkv, dparams, stream);
  if (N <= 32) return launch_bwd_mma<2>(qkv, dout, batch, N, H, pm, nm, inv_freq, qw, kw, qs, eps, dqkv, dparams, stream);
  return launch_bwd_mma<4>(qkv, dout, batch, N, H, pm, nm, inv_freq, qw, kw, qs, eps, dqkv, dparams, stream);
}

}  // namespace tsfmx

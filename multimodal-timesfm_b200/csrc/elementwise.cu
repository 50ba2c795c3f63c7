// Row-wise normalisation kernels that sit between the tcgen05 GEMMs.
//
// One warp owns one row (1280 fp32 for TimesFM 2.5, 768 for Chronos-2): the row is read once with
// 16-byte streaming loads, kept in registers, reduced with warp shuffles and written once.
//
//   rmsnorm               upstream RMSNorm (HF twin modeling_timesfm2_5.py:118-135)
//   norm_residual_norm    `post_ln(a) + x` followed by the next `pre_ln` of a TimesFM 2.5 layer
//                         (HF twin modeling_timesfm2_5.py:378-388), one pass instead of ~8 ATen kernels
#include "common.cuh"

namespace tsfmx {
namespace {

constexpr int WARPS = 8;

template <int OUT>
__device__ __forceinline__ void store_quad(void* out, int64_t row, int cols, int c, float a, float b, float cc,
                                           float d) {
  if constexpr (OUT == TSFMX_DT_F32) {
    st_stream_f4(reinterpret_cast<float*>(out) + row * cols + c, make_float4(a, b, cc, d));
  } else if constexpr (OUT == TSFMX_DT_BF16) {
    st_stream_u2(reinterpret_cast<__nv_bfloat16*>(out) + row * cols + c,
                 make_uint2(pack_bf16x2(a, b), pack_bf16x2(cc, d)));
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * 2 * cols + c;
    uint2 h, l;
    split_bf16x2(a, b, h.x, l.x);
    split_bf16x2(cc, d, h.y, l.y);
    st_stream_u2(o, h);
    st_stream_u2(o + cols, l);
  }
}

// rs = 1 / sqrt(mean(v^2) + eps) over the whole row (all lanes hold the result)
template <int NV>
__device__ __forceinline__ float row_rsqrt_mean_sq(const float4 (&v)[NV], int cols, float eps) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  ss = warp_sum(ss);
  return 1.0f / sqrtf(ss / static_cast<float>(cols) + eps);
}

template <int NV, int OUT>
__global__ void __launch_bounds__(WARPS * 32) rmsnorm_kernel(const float* __restrict__ x, int64_t rows,
                                                             const float* __restrict__ w, float eps, void* out) {
  constexpr int COLS = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * WARPS + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * WARPS) {
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ld_stream_f4(x + r * COLS + 4 * (lane + 32 * j));
    const float rs = row_rsqrt_mean_sq<NV>(v, COLS, eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 4 * (lane + 32 * j);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
      store_quad<OUT>(out, r, COLS, c, ww.x * (v[j].x * rs), ww.y * (v[j].y * rs), ww.z * (v[j].z * rs),
                      ww.w * (v[j].w * rs));
    }
  }
}

template <int NV, int A_BF16, int OUT>
__global__ void __launch_bounds__(WARPS * 32) norm_residual_norm_kernel(
    const void* __restrict__ a, const float* x, int64_t rows, const float* __restrict__ w_post,
    const float* __restrict__ w_next, float eps, float* y, void* yn) {
  constexpr int COLS = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * WARPS + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * WARPS) {
    float4 v[NV], xx[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 4 * (lane + 32 * j);
      if constexpr (A_BF16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a) + r * COLS + c);
        const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        v[j] = make_float4(__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1));
      } else {
        v[j] = ld_stream_f4(reinterpret_cast<const float*>(a) + r * COLS + c);
      }
    }
    // the residual row is fetched NOW, together with a: y may alias x (in-place update), so the compiler would not lift
    // these loads above the stores below by itself, and the row would be read one 16-byte chunk per round trip after the
    // first reduction.  A lane only ever overwrites the elements it has read itself.
#pragma unroll
    for (int j = 0; j < NV; ++j) xx[j] = ld_stream_f4(x + r * COLS + 4 * (lane + 32 * j));
    if (w_post != nullptr) {
      const float rs = row_rsqrt_mean_sq<NV>(v, COLS, eps);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 4 * (lane + 32 * j);
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w_post + c));
        v[j].x = ww.x * (v[j].x * rs), v[j].y = ww.y * (v[j].y * rs);
        v[j].z = ww.z * (v[j].z * rs), v[j].w = ww.w * (v[j].w * rs);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 4 * (lane + 32 * j);
      v[j].x += xx[j].x, v[j].y += xx[j].y, v[j].z += xx[j].z, v[j].w += xx[j].w;
      if (y != nullptr) st_stream_f4(y + r * COLS + c, v[j]);
    }
    if (yn != nullptr) {
      if (w_next != nullptr) {
        const float rs = row_rsqrt_mean_sq<NV>(v, COLS, eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = 4 * (lane + 32 * j);
          const float4 ww = __ldg(reinterpret_cast<const float4*>(w_next + c));
          store_quad<OUT>(yn, r, COLS, c, ww.x * (v[j].x * rs), ww.y * (v[j].y * rs), ww.z * (v[j].z * rs),
                          ww.w * (v[j].w * rs));
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          store_quad<OUT>(yn, r, COLS, 4 * (lane + 32 * j), v[j].x, v[j].y, v[j].z, v[j].w);
      }
    }
  }
}

int grid_for_rows(int64_t rows) {
  const int64_t blocks = (rows + WARPS - 1) / WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8 * 4;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

template <int NV>
int launch_rmsnorm(const float* x, int64_t rows, const float* w, float eps, int out_dtype, void* out,
                   cudaStream_t stream) {
  const int grid = grid_for_rows(rows);
  const dim3 block(WARPS * 32);
  if (out_dtype == TSFMX_DT_F32) rmsnorm_kernel<NV, TSFMX_DT_F32><<<grid, block, 0, stream>>>(x, rows, w, eps, out);
  else if (out_dtype == TSFMX_DT_BF16) rmsnorm_kernel<NV, TSFMX_DT_BF16><<<grid, block, 0, stream>>>(x, rows, w, eps, out);
  else rmsnorm_kernel<NV, TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(x, rows, w, eps, out);
  return check_last_launch("rmsnorm");
}

template <int NV, int A_BF16>
int launch_nrn(const void* a, const float* x, int64_t rows, const float* w_post, const float* w_next, float eps,
               float* y, int yn_dtype, void* yn, cudaStream_t stream) {
  const int grid = grid_for_rows(rows);
  const dim3 block(WARPS * 32);
  if (yn_dtype == TSFMX_DT_F32)
    norm_residual_norm_kernel<NV, A_BF16, TSFMX_DT_F32><<<grid, block, 0, stream>>>(a, x, rows, w_post, w_next, eps, y, yn);
  else if (yn_dtype == TSFMX_DT_BF16)
    norm_residual_norm_kernel<NV, A_BF16, TSFMX_DT_BF16><<<grid, block, 0, stream>>>(a, x, rows, w_post, w_next, eps, y, yn);
  else
    norm_residual_norm_kernel<NV, A_BF16, TSFMX_DT_BF16_SPLIT><<<grid, block, 0, stream>>>(a, x, rows, w_post, w_next, eps, y, yn);
  return check_last_launch("norm_residual_norm");
}

}  // namespace
}  // namespace tsfmx

using namespace tsfmx;

extern "C" int tsfmx_rmsnorm(const float* x, int64_t rows, int32_t cols, const float* w, float eps,
                             int32_t out_dtype, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(x != nullptr && w != nullptr && out != nullptr, "rmsnorm: NULL pointer");
  TSFMX_REQUIRE(rows >= 0, "rmsnorm: bad rows");
  TSFMX_REQUIRE(out_dtype >= TSFMX_DT_F32 && out_dtype <= TSFMX_DT_BF16_SPLIT, "rmsnorm: bad out_dtype");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "rmsnorm: pointers must be 16-byte aligned");
  if (rows == 0) return TSFMX_OK;
  switch (cols) {
    case 1280: return launch_rmsnorm<10>(x, rows, w, eps, out_dtype, out, stream);
    case 768: return launch_rmsnorm<6>(x, rows, w, eps, out_dtype, out, stream);
    default:
      set_error("rmsnorm: cols=%d unsupported (1280 for TimesFM 2.5, 768 for Chronos-2)", cols);
      return TSFMX_ERR_UNSUPPORTED;
  }
}

extern "C" int tsfmx_norm_residual_norm(const void* a, int32_t a_dtype, const float* x, int64_t rows, int32_t cols,
                                        const float* w_post, const float* w_next, float eps, float* y,
                                        int32_t yn_dtype, void* yn, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TSFMX_REQUIRE(a != nullptr && x != nullptr, "norm_residual_norm: NULL pointer");
  TSFMX_REQUIRE(y != nullptr || yn != nullptr, "norm_residual_norm: no output requested");
  TSFMX_REQUIRE(a_dtype == TSFMX_DT_F32 || a_dtype == TSFMX_DT_BF16, "norm_residual_norm: a must be f32 or bf16");
  TSFMX_REQUIRE(yn_dtype >= TSFMX_DT_F32 && yn_dtype <= TSFMX_DT_BF16_SPLIT, "norm_residual_norm: bad yn_dtype");
  TSFMX_REQUIRE(reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(yn) % 16 == 0,
                "norm_residual_norm: pointers must be 16-byte aligned");
  if (rows == 0) return TSFMX_OK;
  const bool abf = a_dtype == TSFMX_DT_BF16;
  switch (cols) {
    case 1280:
      return abf ? launch_nrn<10, 1>(a, x, rows, w_post, w_next, eps, y, yn_dtype, yn, stream)
                 : launch_nrn<10, 0>(a, x, rows, w_post, w_next, eps, y, yn_dtype, yn, stream);
    case 768:
      return abf ? launch_nrn<6, 1>(a, x, rows, w_post, w_next, eps, y, yn_dtype, yn, stream)
                 : launch_nrn<6, 0>(a, x, rows, w_post, w_next, eps, y, yn_dtype, yn, stream);
    default:
      set_error("norm_residual_norm: cols=%d unsupported (1280 for TimesFM 2.5, 768 for Chronos-2)", cols);
      return TSFMX_ERR_UNSUPPORTED;
  }
}

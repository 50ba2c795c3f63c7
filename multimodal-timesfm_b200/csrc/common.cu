// Error state, launch accounting and device probing shared by all entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace tsfmx {

namespace {
thread_local char g_error[1024] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int g_force_simt_attention = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// Called right after every kernel launch: counts it and surfaces launch-time errors.
int check_last_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TSFMX_ERR_CUDA;
  }
  return TSFMX_OK;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tsfmx

extern "C" int tsfmx_abi_version(void) { return TSFMX_ABI_VERSION; }

extern "C" int tsfmx_sizeof_gemm_args(void) { return static_cast<int>(sizeof(tsfmx_gemm_args)); }

extern "C" const char* tsfmx_last_error(void) { return tsfmx::g_error; }

extern "C" uint64_t tsfmx_launch_count(void) { return tsfmx::g_launches.load(std::memory_order_relaxed); }

extern "C" int tsfmx_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    tsfmx::set_error("no CUDA device available (%s); tsfmx_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return TSFMX_ERR_NO_DEVICE;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    tsfmx::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    return TSFMX_ERR_CUDA;
  }
  if (prop.major != 10) {
    tsfmx::set_error("device %d is sm_%d%d; tsfmx_b200 kernels are built for sm_100a only", device, prop.major,
                     prop.minor);
    return TSFMX_ERR_NO_DEVICE;
  }
  return TSFMX_OK;
}

// common.cuh -- shared host/device helpers for libsparsifyme_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/spfy_b200.h"

namespace spfy {

// ---- error reporting: thread-local message, int status --------------------
inline char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

#define SPFY_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::spfy::fail(SPFY_E_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                          cudaGetErrorString(e__));                                         \
  } while (0)

// call after every kernel launch
#define SPFY_LAUNCH_OK(name)                                                                \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return ::spfy::fail(SPFY_E_CUDA, "launch of %s failed: %s", name,                     \
                          cudaGetErrorString(e__));                                         \
    ::spfy::launch_counter().fetch_add(1, std::memory_order_relaxed);                       \
  } while (0)

// Development switches (ring depth, debug modes that skip work, kernel-choice overrides) are read from the
// environment ONLY in builds made with -DSPFY_DEV_SWITCHES (build.py --dev); the shipped library ignores them,
// so no environment variable can change results or skip work in production.
inline const char* dev_switch(const char* name) {
#ifdef SPFY_DEV_SWITCHES
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return ceil_div(a, b) * b; }

// device properties are queried once per device and cached
struct DeviceInfo {
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int max_smem_optin = 0;
};
inline int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  static std::atomic<int> ready[64];
  int dev = 0;
  SPFY_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPFY_E_CUDA, "device index %d out of range", dev);
  if (!ready[dev].load(std::memory_order_acquire)) {
    DeviceInfo d;
    SPFY_CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    SPFY_CUDA_OK(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    SPFY_CUDA_OK(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    SPFY_CUDA_OK(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev] = d;
    ready[dev].store(1, std::memory_order_release);
  }
  *out = cache[dev];
  return SPFY_OK;
}

// CUDA loads device code lazily, kernel by kernel, at first use (tens of milliseconds for the big ones).
// Each translation unit lists its kernels here so that spfy_init() can pay that once, outside anybody's timers.
template <typename K>
inline void touch_kernel(K kernel) {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, kernel);
}
void warm_prune_kernels();
void warm_spmma_kernels();
void warm_spmm_kernels();
void warm_gemm_kernels();

inline size_t dtype_bytes(int dtype) {
  switch (dtype) {
    case SPFY_F16:
    case SPFY_BF16: return 2;
    case SPFY_F32: return 4;
    case SPFY_F64: return 8;
    default: return 0;
  }
}

}  // namespace spfy

// api.cu -- library-level entry points of libsparsifyme_b200.so
#include "common.cuh"

namespace spfy {
namespace {

template <typename S, typename D>
__device__ __forceinline__ D cvt(S v);
template <> __device__ __forceinline__ __half cvt<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float cvt<__half, float>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float cvt<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }

// streaming element-wise conversion: 8 elements per thread per trip when both sides are 16-byte
// aligned, scalar otherwise / for the tail
template <typename S, typename D>
__global__ void __launch_bounds__(256)
convert_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n, int vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (vec) {
    const size_t chunks = n / 8;
    for (size_t c = tid; c < chunks; c += nthreads) {
      S in[8];
      D out[8];
      if (sizeof(S) == 4) {
        *reinterpret_cast<uint4*>(in) = *reinterpret_cast<const uint4*>(src + c * 8);
        *reinterpret_cast<uint4*>(in + 4) = *reinterpret_cast<const uint4*>(src + c * 8 + 4);
      } else {
        *reinterpret_cast<uint4*>(in) = *reinterpret_cast<const uint4*>(src + c * 8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = cvt<S, D>(in[i]);
      if (sizeof(D) == 4) {
        *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<uint4*>(out);
        *reinterpret_cast<uint4*>(dst + c * 8 + 4) = *reinterpret_cast<uint4*>(out + 4);
      } else {
        *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<uint4*>(out);
      }
    }
    done = chunks * 8;
  }
  for (size_t i = done + tid; i < n; i += nthreads) dst[i] = cvt<S, D>(src[i]);
}

template <typename S, typename D>
int launch_convert(const void* src, void* dst, size_t n, cudaStream_t s) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  size_t blocks = ceil_div(ceil_div(n, 8), 256);
  const size_t cap = (size_t)di.sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (!blocks) blocks = 1;
  const int vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  convert_kernel<S, D><<<(unsigned)blocks, 256, 0, s>>>((const S*)src, (D*)dst, n, vec);
  SPFY_LAUNCH_OK("convert_kernel");
  return SPFY_OK;
}

}  // namespace
}  // namespace spfy

using namespace spfy;

extern "C" {

int spfy_version(void) { return 100; }  // 0.1.0

int spfy_init(void) {
  static std::atomic<int> done[64];
  int dev = 0;
  SPFY_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPFY_E_CUDA, "device index %d out of range", dev);
  if (done[dev].load(std::memory_order_acquire)) return SPFY_OK;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  spfy::warm_prune_kernels();
  spfy::warm_spmma_kernels();
  spfy::warm_spmm_kernels();
  touch_kernel(convert_kernel<float, __half>);
  touch_kernel(convert_kernel<float, __nv_bfloat16>);
  touch_kernel(convert_kernel<__half, float>);
  touch_kernel(convert_kernel<__nv_bfloat16, float>);
  (void)cudaGetLastError();
  done[dev].store(1, std::memory_order_release);
  return SPFY_OK;
}

const char* spfy_last_error_string(void) { return spfy::err_buf(); }

uint64_t spfy_launch_count(void) { return spfy::launch_counter().load(std::memory_order_relaxed); }

int spfy_convert(int src_dtype, int dst_dtype, const void* src, void* dst, size_t count,
                 spfy_stream_t stream) {
  if (count == 0) return SPFY_OK;
  if (!src || !dst) return fail(SPFY_E_INVALID, "convert: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (src_dtype == SPFY_F32 && dst_dtype == SPFY_F16) return launch_convert<float, __half>(src, dst, count, s);
  if (src_dtype == SPFY_F32 && dst_dtype == SPFY_BF16) return launch_convert<float, __nv_bfloat16>(src, dst, count, s);
  if (src_dtype == SPFY_F16 && dst_dtype == SPFY_F32) return launch_convert<__half, float>(src, dst, count, s);
  if (src_dtype == SPFY_BF16 && dst_dtype == SPFY_F32) return launch_convert<__nv_bfloat16, float>(src, dst, count, s);
  return fail(SPFY_E_UNSUPPORTED, "convert: %d -> %d is not supported (F32 <-> F16/BF16 only)", src_dtype, dst_dtype);
}

}  // extern "C"

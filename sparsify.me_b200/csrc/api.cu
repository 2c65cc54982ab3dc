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

// ---------------------------------------------------------------- pruned-layer container (host memory only)
namespace {
struct PackedHeader {
  char magic[8];
  uint32_t version;
  int32_t dtype, layout;
  uint32_t reserved;
  uint64_t rows, cols, vals_bytes, meta_bytes, checksum;
};
static_assert(sizeof(PackedHeader) == SPFY_PACKED_HEADER_BYTES, "container header is 64 bytes");
const char kPackedMagic[8] = {'S', 'P', 'F', 'Y', '2', '4', 0, 0};

uint64_t fnv1a64(const uint8_t* p, size_t n, uint64_t h) {
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}
}  // namespace

int spfy_packed_bytes(int dtype, size_t rows, size_t cols, int layout, size_t* bytes) {
  size_t vb = 0, mb = 0;
  int rc = spfy_compressed_bytes(dtype, rows, cols, layout, &vb, &mb);
  if (rc) return rc;
  if (bytes) *bytes = SPFY_PACKED_HEADER_BYTES + vb + mb;
  return SPFY_OK;
}

int spfy_packed_write(int dtype, int layout, size_t rows, size_t cols, const void* host_vals,
                      const void* host_meta, void* dst, size_t dst_bytes) {
  size_t vb = 0, mb = 0;
  int rc = spfy_compressed_bytes(dtype, rows, cols, layout, &vb, &mb);
  if (rc) return rc;
  if (!dst || (vb && !host_vals) || (mb && !host_meta)) return fail(SPFY_E_INVALID, "packed_write: null pointer");
  if (dst_bytes < SPFY_PACKED_HEADER_BYTES + vb + mb)
    return fail(SPFY_E_WORKSPACE, "packed_write: buffer %zu < %zu bytes", dst_bytes, (size_t)SPFY_PACKED_HEADER_BYTES + vb + mb);
  PackedHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kPackedMagic, 8);
  h.version = 1;
  h.dtype = dtype;
  h.layout = layout;
  h.rows = rows;
  h.cols = cols;
  h.vals_bytes = vb;
  h.meta_bytes = mb;
  h.checksum = fnv1a64((const uint8_t*)host_meta, mb, fnv1a64((const uint8_t*)host_vals, vb, 0xcbf29ce484222325ull));
  uint8_t* out = (uint8_t*)dst;
  memcpy(out, &h, sizeof(h));
  memcpy(out + sizeof(h), host_vals, vb);
  memcpy(out + sizeof(h) + vb, host_meta, mb);
  return SPFY_OK;
}

int spfy_packed_read(const void* src, size_t src_bytes, int* dtype, int* layout, size_t* rows, size_t* cols,
                     size_t* vals_offset, size_t* vals_bytes, size_t* meta_offset, size_t* meta_bytes) {
  if (!src || src_bytes < SPFY_PACKED_HEADER_BYTES) return fail(SPFY_E_INVALID, "packed_read: buffer shorter than a header");
  PackedHeader h;
  memcpy(&h, src, sizeof(h));
  if (memcmp(h.magic, kPackedMagic, 8) != 0) return fail(SPFY_E_INVALID, "packed_read: bad magic");
  if (h.version != 1) return fail(SPFY_E_UNSUPPORTED, "packed_read: container version %u", h.version);
  size_t vb = 0, mb = 0;
  int rc = spfy_compressed_bytes(h.dtype, (size_t)h.rows, (size_t)h.cols, h.layout, &vb, &mb);
  if (rc) return rc;
  if (vb != h.vals_bytes || mb != h.meta_bytes)
    return fail(SPFY_E_INVALID, "packed_read: sizes in the header do not match the shape");
  if (src_bytes < SPFY_PACKED_HEADER_BYTES + vb + mb) return fail(SPFY_E_INVALID, "packed_read: truncated payload");
  const uint8_t* p = (const uint8_t*)src + SPFY_PACKED_HEADER_BYTES;
  if (fnv1a64(p + vb, mb, fnv1a64(p, vb, 0xcbf29ce484222325ull)) != h.checksum)
    return fail(SPFY_E_INVALID, "packed_read: checksum mismatch");
  if (dtype) *dtype = h.dtype;
  if (layout) *layout = h.layout;
  if (rows) *rows = (size_t)h.rows;
  if (cols) *cols = (size_t)h.cols;
  if (vals_offset) *vals_offset = SPFY_PACKED_HEADER_BYTES;
  if (vals_bytes) *vals_bytes = vb;
  if (meta_offset) *meta_offset = SPFY_PACKED_HEADER_BYTES + vb;
  if (meta_bytes) *meta_bytes = mb;
  return SPFY_OK;
}

}  // extern "C"

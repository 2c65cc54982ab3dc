// write_probe.cu -- what this GPU's DRAM takes as pure writes, pure reads and a mix, with plain kernels of our own
// (context for the write-heavy spmma classes: their bound is the write rate, not the copy rate).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/write_probe tools/write_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void write_kernel(uint4* p, size_t n, int mode) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  if (mode < 2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      if (mode == 0) p[i] = v;
      else asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  } else {
    // 256-bit stores (sm_100): plain, and with the L2 evict-first policy (only the 256-bit form takes it)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (size_t)gridDim.x * blockDim.x) {
      if (mode == 2)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(p + 2 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      else
        asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(p + 2 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
}
// each CTA writes a contiguous chunk (DRAM page locality) instead of a grid-strided interleave
__global__ void write_chunk_kernel(uint4* p, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  const size_t per = (n + gridDim.x - 1) / gridDim.x, b = per * blockIdx.x, e = b + per < n ? b + per : n;
  for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) p[i] = v;
}
__global__ void read_kernel(const uint4* p, size_t n, uint32_t* sink) {
  uint32_t acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
// r reads per w writes (in units of 16 bytes per thread trip): the mix of a GEMM that reads B and writes D
__global__ void mix_kernel(const uint4* src, uint4* dst, size_t n, int reads, uint32_t* sink) {
  uint32_t acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    for (int r = 0; r < reads; ++r) {
      const uint4 v = src[i + (size_t)r * n];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    dst[i] = make_uint4(acc, 2, 3, 4);
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static double time_ms(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)1 << 30, n = bytes / 16;
  uint4 *a, *b; uint32_t* sink;
  cudaMalloc(&a, 4 * bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, 4 * bytes);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const char* names[4] = {"st.global.v4", "st.global.cs.v4", "st.global.v8.b32", "st.global.L2::evict_first.v8.b32"};
  for (int mode = 0; mode < 4; ++mode)
    for (int ctas : {sms * 2, sms * 8, sms * 32}) {
      const double ms = time_ms([&] { write_kernel<<<ctas, 512>>>(b, n, mode); });
      printf("write 1 GiB  %-30s grid %5d x 512: %7.0f GB/s\n", names[mode], ctas, bytes / ms / 1e6);
    }
  for (int ctas : {sms, sms * 8}) {
    const double ms = time_ms([&] { write_chunk_kernel<<<ctas, 512>>>(b, n); });
    printf("write 1 GiB  contiguous chunk per CTA       grid %5d x 512: %7.0f GB/s\n", ctas, bytes / ms / 1e6);
  }
  { const double ms = time_ms([&] { cudaMemsetAsync(b, 0, bytes); }); printf("cudaMemset 1 GiB: %7.0f GB/s\n", bytes / ms / 1e6); }
  { const double ms = time_ms([&] { read_kernel<<<sms * 8, 512>>>(a, n, sink); }); printf("read 1 GiB: %7.0f GB/s\n", bytes / ms / 1e6); }
  { const double ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D 1 GiB: %7.0f GB/s (read + write)\n", 2 * bytes / ms / 1e6); }
  for (int reads : {1, 2, 3, 4}) {
    const size_t nn = n / 4;  // dst 256 MiB, src reads x 256 MiB
    const double ms = time_ms([&] { mix_kernel<<<sms * 8, 512>>>(a, b, nn, reads, sink); });
    printf("mix %d read : 1 write (16 B each), %4zu MiB written: %7.0f GB/s total, %7.0f GB/s of writes\n", reads, nn * 16 >> 20,
           (reads + 1) * nn * 16 / ms / 1e6, nn * 16 / ms / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

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

// The store pattern of a GEMM epilogue on a row-major [rows x cols] output (16-bit elements): a CTA owns tiles of
// `tile_rows` rows x `tile_bytes` bytes, consecutive CTAs take consecutive column tiles (round robin), every warp store
// instruction covers 512 bytes = (512 / tile_bytes) rows of the tile.  Optionally every tile is preceded by reading
// `read_rows` rows x tile_bytes of a second matrix of the same pitch (the B tile of a small-K GEMM).
__global__ void tile_write_kernel(uint4* out, const uint4* in, size_t pitch_bytes, uint32_t rows, uint32_t tile_rows,
                                  uint32_t tile_bytes, uint32_t col_tiles, uint32_t read_rows, uint32_t* sink) {
  const uint32_t row_tiles = rows / tile_rows, total = row_tiles * col_tiles;
  const uint32_t per_row = tile_bytes / 16, rows_per_pass = blockDim.x / per_row;
  const uint32_t r_in = threadIdx.x / per_row, c_in = threadIdx.x % per_row;
  uint32_t acc = 0;
  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    const uint32_t ct = t / row_tiles, rt = t % row_tiles;  // row tile fastest, like the m-groups of a unit list
    const size_t col0 = (size_t)ct * tile_bytes;
    for (uint32_t r = r_in; r < read_rows; r += rows_per_pass) {
      const uint4 v = in[((size_t)r * pitch_bytes + col0) / 16 + c_in];
      acc ^= v.x ^ v.w;
    }
    const uint4 v = make_uint4(acc, 2, 3, 4);
    for (uint32_t r = r_in; r < tile_rows; r += rows_per_pass)
      out[((size_t)(rt * tile_rows + r) * pitch_bytes + col0) / 16 + c_in] = v;
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
  {
    // D of 256 x 64 x 401408 (k <= 64 class): 256 rows x 802816 bytes; B: 64 rows of the same pitch
    const size_t pitch = 401408 * 2;
    uint4 *d2, *b2;
    cudaMalloc(&d2, 256 * pitch); cudaMalloc(&b2, 64 * pitch);
    cudaMemset(b2, 1, 64 * pitch);
    for (int per_sm : {1, 2, 4})
      for (uint32_t tile_bytes : {256u, 512u, 1024u})
        for (uint32_t read_rows : {0u, 64u}) {
          const uint32_t col_tiles = (uint32_t)(pitch / tile_bytes);
          const double ms = time_ms([&] { tile_write_kernel<<<sms * per_sm, 512>>>(d2, b2, pitch, 256, 256, tile_bytes, col_tiles, read_rows, sink); });
          const double wb = 256.0 * pitch, rb = (double)read_rows * pitch;
          printf("epilogue pattern: %d CTA/SM, tiles of 256 rows x %4u B, pitch %zu B, %2u rows read per tile: %7.0f GB/s total (%6.1f us)\n",
                 per_sm, tile_bytes, pitch, read_rows, (wb + rb) / ms / 1e6, ms * 1e3);
        }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

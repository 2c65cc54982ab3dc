// tmem_probe.cu -- how fast tensor memory can be read back (tcgen05.ld), per SM: the epilogue of every GEMM tile moves
// its 128 x 128 fp32 accumulator (64 KiB) through this path, and for the write-heavy spmma classes (k <= 64) that is the
// longest stage of a tile.  One CTA per SM, `warps` warps (4 = one per lane quarter, 8 = two per quarter on different
// column halves, like the spmma epilogue); every warp reads `cols` columns of its 32 lanes `iters` times.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Isparsify.me_b200/csrc -Iinclude -o tools/bin/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace spfy::ptx;

template <int X>  // X = 32: two .x32 loads per wait; 16: four .x16 loads per wait
__global__ void __launch_bounds__(256, 1) ld_kernel(int iters, int cols_per_warp, uint32_t* sink, long long* cycles) {
  __shared__ uint32_t tmem_slot;
  const uint32_t warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot;
  const uint32_t quarter = warp & 3u, half = warp >> 2;
  const uint32_t col0 = half * (512u / (warps > 4 ? 2u : 1u));
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols_per_warp; c += 64) {
      uint32_t r[64];
      const uint32_t taddr = base + ((quarter * 32u) << 16) + col0 + (uint32_t)((it * 64 + c) % (int)(512u / (warps > 4 ? 2u : 1u)));
      if (X == 32) {
        tmem_ld_x32(taddr, r);
        tmem_ld_x32(taddr + 32, r + 32);
      } else {
        tmem_ld_x16(taddr, r);
        tmem_ld_x16(taddr + 16, r + 16);
        tmem_ld_x16(taddr + 32, r + 32);
        tmem_ld_x16(taddr + 48, r + 48);
      }
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 64; i += 16) acc ^= r[i];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x9e3779b9u) *sink = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  uint32_t* sink;
  long long* cyc;
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount);
  long long h[256];
  printf("warps,shape,bytes_per_sm,cycles,B_per_clk_per_SM\n");
  for (int warps : {1, 4, 8}) {
    for (int x : {32, 16}) {
      const int iters = 2000, cols = 128;  // every warp: 32 lanes x 128 columns x 4 B = 16 KiB per iteration
      for (int rep = 0; rep < 2; ++rep) {
        if (x == 32) ld_kernel<32><<<p.multiProcessorCount, warps * 32>>>(iters, cols, sink, cyc);
        else ld_kernel<16><<<p.multiProcessorCount, warps * 32>>>(iters, cols, sink, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, cyc, sizeof(long long) * p.multiProcessorCount, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < p.multiProcessorCount; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)warps * iters * cols * 32 * 4;
      printf("%d,32x32b.x%d,%.0f,%lld,%.1f\n", warps, x, bytes, mx, bytes / mx);
    }
  }
  return 0;
}

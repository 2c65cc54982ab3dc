// spmma_sm100.cu -- 2:4 structured-sparse GEMM on tcgen05.mma.sp (sm_100a)
//
// Replaces cusparseLtMatmul (reference: include/sparsify.me/spmma.hxx:106-114):
//     D[m x n] = alpha * A(2:4)[m x k] * op(B)[k x n] + beta * C[m x n]
// all row-major like the reference's descriptors (spmma.hxx:56-64), fp16 or bf16
// in, fp32 accumulate in TMEM, fp16/bf16 out.
//
// Shape of the kernel (one persistent CTA per SM, 256 threads, 1 CTA/SM):
//   warp 0   producer   : cp.async.bulk of the pre-swizzled A-value tile (<=16 KiB)
//                         and its metadata tile (<=2 KiB), TMA tensor loads of the
//                         B tile, all landing on one mbarrier per stage
//   warp 1   MMA issuer : tcgen05.cp (metadata smem -> TMEM), then up to four
//                         tcgen05.mma.sp.kind::f16 (M128 x N128 x K32) per stage;
//                         tcgen05.commit releases the stage / publishes the accumulator
//   warp 2   TMEM allocator (512 columns: 2 x 128 accumulator + 2 x 4 metadata)
//   warps 4-7 epilogue  : tcgen05.ld -> alpha/beta -> fp16/bf16 -> 128B-swizzled smem
//                         -> TMA tensor store (clips the ragged M / N edges)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the
// main loop of tile i+1.  Every ResNet shape in datasets/*.csv is HBM-bound for this
// operator, so the design goal is: read B exactly once from HBM, keep
// many bytes in flight, never stall the stream on the epilogue.
//
// The compressed operand comes from spfy_prune24(layout = SPFY_LAYOUT_SM100): tile
// (mt,kt) of A covers rows [128mt,128mt+128) x logical cols [128kt,128kt+128); its
// value tile is the exact 128B-swizzled K-major shared-memory image (row r at r*128,
// 16-byte chunk c stored at chunk c ^ (r & 7)), its metadata tile the exact
// `tcgen05.cp.128x128b` source image of the kind::f16 sparse-metadata TMEM layout.
#include "common.cuh"

#include <cstdlib>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)

namespace spfy {
namespace {

constexpr int BM = 128;        // rows of A per tile (UMMA M)
constexpr int BN = 128;        // columns of B/D per tile (UMMA N)
constexpr int BK = 128;        // logical K per stage (4 MMAs of K=32)
constexpr int STAGES = 3;
constexpr int A_TILE_BYTES = 16384;
constexpr int E_TILE_BYTES = 2048;
constexpr int B_STAGE_BYTES = BK * BN * 2;
constexpr int C_BUF_BYTES = BM * 64 * 2;
constexpr int C_BUFS = 2;
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int TMEM_E_COL = 2 * BN;  // metadata columns start after the two accumulators

constexpr int SMEM_A = 0;
constexpr int SMEM_B = SMEM_A + STAGES * A_TILE_BYTES;
constexpr int SMEM_C = SMEM_B + STAGES * B_STAGE_BYTES;
constexpr int SMEM_E = SMEM_C + C_BUFS * C_BUF_BYTES;
constexpr int SMEM_BAR = SMEM_E + STAGES * E_TILE_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_TMEM_PTR = SMEM_BAR + NUM_BARS * 8;
constexpr int SMEM_TOTAL = SMEM_TMEM_PTR + 16;
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack for the 1024-byte alignment

constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;
constexpr uint64_t HINT_EVICT_NORMAL = 0x1000000000000000ull;

struct SpmmaParams {
  const uint8_t* a_vals;
  const uint8_t* a_meta;
  const void* C;  // only read when beta != 0
  size_t ldc;
  uint32_t m, n, k;
  uint32_t m_tiles, n_tiles, k_tiles;
  float alpha, beta;
  uint32_t idesc;
  uint64_t hint_b;
  uint32_t dbg;  // development switches (SPFY_SPMMA_DEBUG), 0 in production
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must become a trap, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > 4000000000ull) {
      printf("spfy spmma: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// metadata: 128 lanes x 128 bits, shared memory -> TMEM (4 columns)
__device__ __forceinline__ void tc_cp_128x128b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// D[tmem] (+)= A(2:4)[smem] * B[smem], metadata in TMEM
__device__ __forceinline__ void tc_mma_sp_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t tmem_e, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.sp.cta_group::1.kind::f16 [%0], %1, %2, [%3], %4, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(tmem_e), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// shared-memory matrix descriptor (sm_100 format, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16 |
         (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32 | (uint64_t)1 << 46 |
         (uint64_t)layout_type << 61;
}
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_NONE = 0;

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  if (BF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&w));
  }
}

// ------------------------------------------------------------------- kernel
template <bool BF16, bool OPB_T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
spmma_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_d,
             const SpmmaParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = smem_base + SMEM_BAR;              // [STAGES]
  const uint32_t bar_empty = bar_full + STAGES * 8;            // [STAGES]
  const uint32_t bar_acc_full = bar_empty + STAGES * 8;        // [2]
  const uint32_t bar_acc_empty = bar_acc_full + 2 * 8;         // [2]
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + SMEM_TMEM_PTR);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + a * 8, 1);
      mbar_init(bar_acc_empty + a * 8, 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_base + SMEM_TMEM_PTR, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t num_tiles = P.m_tiles * P.n_tiles;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint32_t m_blk = t % P.m_tiles, n_blk = t / P.m_tiles;
        const uint32_t rows_left = P.m - m_blk * BM;
        const uint32_t rows_valid = rows_left >= BM ? BM : ((rows_left + 15u) & ~15u);
        const uint32_t tx = rows_valid * 128u + rows_valid * 16u + ((P.dbg & 4) ? 0u : B_STAGE_BYTES);
        for (uint32_t kt = 0; kt < P.k_tiles; ++kt) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);
          const uint32_t full = bar_full + stage * 8;
          mbar_expect_tx(full, tx);
          const size_t tile = (size_t)m_blk * P.k_tiles + kt;
          bulk_load_1d(smem_base + SMEM_A + stage * A_TILE_BYTES, P.a_vals + tile * A_TILE_BYTES,
                       rows_valid * 128u, full, HINT_EVICT_LAST);
          bulk_load_1d(smem_base + SMEM_E + stage * E_TILE_BYTES, P.a_meta + tile * E_TILE_BYTES,
                       rows_valid * 16u, full, HINT_EVICT_LAST);
          const uint32_t sb = smem_base + SMEM_B + stage * B_STAGE_BYTES;
          if (P.dbg & 4) {
          } else if (!OPB_T) {
            // B is k x n row-major: boxes of [BK rows of k][64 columns of n] -> MN-major SW128
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &tmap_b, (int)(n_blk * BN + j * 64), (int)(kt * BK),
                          full, P.hint_b);
          } else {
            // B is n x k row-major: boxes of [BN rows of n][64 columns of k] -> K-major SW128
#pragma unroll
            for (int j = 0; j < BK / 64; ++j)
              tma_load_2d(sb + j * (BN * 128), &tmap_b, (int)(kt * BK + j * 64), (int)(n_blk * BN),
                          full, P.hint_b);
          }
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0, kiter = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(bar_acc_empty + as * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (uint32_t kt = 0; kt < P.k_tiles; ++kt, ++kiter) {
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          const uint32_t e_col = tmem_base + TMEM_E_COL + (kiter & 1) * 4;
          tc_cp_128x128b(e_col, make_smem_desc(smem_base + SMEM_E + stage * E_TILE_BYTES, 16, 128, LAYOUT_NONE));
          const uint32_t k_left = P.k - kt * BK;
          const uint32_t nk = k_left >= BK ? 4u : (k_left + 31u) / 32u;
          const uint32_t sa = smem_base + SMEM_A + stage * A_TILE_BYTES;
          const uint32_t sb = smem_base + SMEM_B + stage * B_STAGE_BYTES;
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j) {
            if (j < nk && !(P.dbg & 8)) {
              // A: K-major SW128, 32 logical = 16 stored halves = 32 bytes per MMA
              const uint64_t da = make_smem_desc(sa + j * 32, 16, 1024, LAYOUT_SW128);
              uint64_t db;
              if (!OPB_T)  // MN-major SW128: 8 k-rows per 1024B atom, 64-column groups BK*128 apart
                db = make_smem_desc(sb + j * 32 * 128, BK * 128, 1024, LAYOUT_SW128);
              else         // K-major SW128: two 64-wide k halves, 64 bytes per MMA inside a row
                db = make_smem_desc(sb + (j >> 1) * (BN * 128) + (j & 1) * 64, 16, 1024, LAYOUT_SW128);
              const uint32_t col = e_col + j;
              tc_mma_sp_f16(tmem_d, da, db, col & ~1u, P.idesc | (col & 1u), (kt | j) != 0);
            }
          }
          tc_commit(bar_empty + stage * 8);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        tc_commit(bar_acc_full + as * 8);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const uint32_t ew = warp - 4;              // == warp % 4: TMEM lane quarter
    const uint32_t row = ew * 32 + lane;       // row inside the tile == TMEM lane
    const uint32_t ethread = threadIdx.x - 128;
    uint32_t it = 0, cbuf = 0;
    for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t m_blk = t % P.m_tiles, n_blk = t / P.m_tiles;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const uint32_t m0 = m_blk * BM, n0 = n_blk * BN;
      const bool warp_has_rows = m0 + ew * 32 < P.m;
      mbar_wait(bar_acc_full + as * 8, aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t acc[64];
        if (warp_has_rows && !(P.dbg & 2)) {
          const uint32_t taddr = tmem_base + as * BN + c * 64 + ((ew * 32) << 16);
          tmem_ld_x32(taddr, acc);
          tmem_ld_x32(taddr + 32, acc + 32);
          tmem_wait_ld();
        }
        if (c == BN / 64 - 1) {  // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + as * 8);
        }
        if (ethread == 0) bulk_wait_read<C_BUFS - 1>();  // the buffer we are about to fill is free
        if (!(P.dbg & 32)) epi_bar_sync();
        const uint32_t sc = smem_base + SMEM_C + cbuf * C_BUF_BYTES;
        if (warp_has_rows && !(P.dbg & (2 | 16))) {
          const uint32_t grow = m0 + row;
          const uint32_t gcol0 = n0 + c * 64;
          const bool use_c = P.beta != 0.f && grow < P.m;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = P.alpha * __uint_as_float(acc[q * 8 + e]);
            if (use_c && gcol0 + q * 8 < P.n) {
              const uint4 cw = *reinterpret_cast<const uint4*>(
                  reinterpret_cast<const uint16_t*>(P.C) + (size_t)grow * P.ldc + gcol0 + q * 8);
              const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack2<BF16>(cws[e]);
                v[2 * e] += P.beta * f.x;
                v[2 * e + 1] += P.beta * f.y;
              }
            }
            st_shared_v4(sc + row * 128 + ((q ^ (row & 7)) << 4), pack2<BF16>(v[0], v[1]),
                         pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7]));
          }
        }
        fence_proxy_async_smem();
        if (!(P.dbg & 32)) epi_bar_sync();
        if (ethread == 0 && !(P.dbg & 3)) {
          tma_store_2d(&tmap_d, sc, (int)(n0 + c * 64), (int)m0);
          bulk_commit();
        }
        cbuf ^= 1;
      }
    }
    if (ethread == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encoder(EncodeTiledFn* out) {
  static std::atomic<void*> cached{nullptr};
  void* fn = cached.load(std::memory_order_acquire);
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    SPFY_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn)
      return fail(SPFY_E_CUDA, "cuTensorMapEncodeTiled not available from this driver");
    cached.store(fn, std::memory_order_release);
  }
  *out = (EncodeTiledFn)fn;
  return SPFY_OK;
}

// 2-D row-major tensor [outer x inner] of 16-bit elements, pitch ld elements
int make_tmap_2d(CUtensorMap* map, int dtype, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc;
  int rc = get_encoder(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == SPFY_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SPFY_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %llu ld %llu", (int)r,
                (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)ld);
  return SPFY_OK;
}

template <bool BF16, bool OPB_T>
int launch(const CUtensorMap& tb, const CUtensorMap& td, const SpmmaParams& P, int grid,
           cudaStream_t s) {
  static std::atomic<int> attr_set[64];
  int dev = 0;
  SPFY_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63].load()) {
    SPFY_CUDA_OK(cudaFuncSetAttribute(spmma_kernel<BF16, OPB_T>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    attr_set[dev & 63].store(1);
  }
  spmma_kernel<BF16, OPB_T><<<grid, NUM_THREADS, SMEM_ALLOC, s>>>(tb, td, P);
  SPFY_LAUNCH_OK("spmma_kernel");
  return SPFY_OK;
}

}  // namespace
}  // namespace spfy

using namespace spfy;

extern "C" {

int spfy_spmma_workspace_bytes(int dtype, size_t m, size_t n, size_t k, size_t* bytes) {
  (void)dtype; (void)m; (void)n; (void)k;
  if (bytes) *bytes = 0;  // the kernel needs no global scratch
  return SPFY_OK;
}

int spfy_spmma(int dtype, int opB, size_t m, size_t n, size_t k, float alpha, const void* comp_vals,
               const void* meta, const void* B, size_t ldb, float beta, const void* C, size_t ldc,
               void* D, size_t ldd, void* workspace, size_t workspace_bytes, spfy_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  if (dtype != SPFY_F16 && dtype != SPFY_BF16)
    return fail(SPFY_E_UNSUPPORTED, "spmma: dtype %d (need F16/BF16)", dtype);
  if (opB != SPFY_OP_N && opB != SPFY_OP_T) return fail(SPFY_E_INVALID, "spmma: bad opB %d", opB);
  if (m == 0 || n == 0) return SPFY_OK;
  if (!comp_vals || !meta || !B || !D) return fail(SPFY_E_INVALID, "spmma: null operand");
  if (beta != 0.f && !C) return fail(SPFY_E_INVALID, "spmma: beta != 0 needs C");
  if (k == 0) return fail(SPFY_E_UNSUPPORTED, "spmma: k == 0");
  if (m >= (1u << 31) || n >= (1u << 31) || k >= (1u << 31))
    return fail(SPFY_E_UNSUPPORTED, "spmma: dimension too large");
  const size_t b_inner = opB == SPFY_OP_N ? n : k;
  if (ldb < b_inner || ldd < n || (beta != 0.f && ldc < n))
    return fail(SPFY_E_INVALID, "spmma: leading dimension too small");
  // TMA contract == the reference's own fp16 contract (spmma.hxx:45-49): multiples of 8
  if (ldb % 8 || ldd % 8 || (beta != 0.f && ldc % 8) || n % 8 || (opB == SPFY_OP_T && k % 8))
    return fail(SPFY_E_UNSUPPORTED,
                "spmma: n, ldb, ldc, ldd (and k for opB=T) must be multiples of 8 elements "
                "(n=%zu k=%zu ldb=%zu ldc=%zu ldd=%zu)", n, k, ldb, ldc, ldd);
  if ((uintptr_t)B % 16 || (uintptr_t)D % 16 || (beta != 0.f && (uintptr_t)C % 16) ||
      (uintptr_t)comp_vals % 16 || (uintptr_t)meta % 16)
    return fail(SPFY_E_INVALID, "spmma: operands must be 16-byte aligned");

  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (di.cc_major != 10)
    return fail(SPFY_E_UNSUPPORTED, "spmma: needs an sm_100a device, found sm_%d%d", di.cc_major, di.cc_minor);

  CUtensorMap tb, td;
  if (opB == SPFY_OP_N)
    rc = make_tmap_2d(&tb, dtype, B, n, k, ldb, 64, BK);
  else
    rc = make_tmap_2d(&tb, dtype, B, k, n, ldb, 64, BN);
  if (rc) return rc;
  rc = make_tmap_2d(&td, dtype, D, n, m, ldd, 64, BM);
  if (rc) return rc;

  SpmmaParams P;
  memset(&P, 0, sizeof(P));
  P.a_vals = (const uint8_t*)comp_vals;
  P.a_meta = (const uint8_t*)meta;
  P.C = C;
  P.ldc = ldc;
  P.m = (uint32_t)m;
  P.n = (uint32_t)n;
  P.k = (uint32_t)k;
  P.m_tiles = (uint32_t)ceil_div(m, BM);
  P.n_tiles = (uint32_t)ceil_div(n, BN);
  P.k_tiles = (uint32_t)ceil_div(k, BK);
  P.alpha = alpha;
  P.beta = beta;
  // instruction descriptor (kind::f16, sparse): c=F32, a/b format, b major, N>>3, M>>4
  const uint32_t fmt = dtype == SPFY_BF16 ? 1u : 0u;
  P.idesc = (1u << 2) | (1u << 4) | (fmt << 7) | (fmt << 10) |
            ((opB == SPFY_OP_N ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
            ((uint32_t)(BM >> 4) << 24);
  // B is streamed once when a single row of tiles covers M; otherwise the other
  // m-tiles of the same columns will want it from L2 again.
  P.hint_b = P.m_tiles == 1 ? HINT_EVICT_FIRST : HINT_EVICT_NORMAL;
  {
    const char* e = getenv("SPFY_SPMMA_DEBUG");
    P.dbg = e ? (uint32_t)atoi(e) : 0u;
  }

  const size_t tiles = (size_t)P.m_tiles * P.n_tiles;
  const int grid = (int)(tiles < (size_t)di.sm_count ? tiles : (size_t)di.sm_count);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SPFY_BF16)
    return opB == SPFY_OP_N ? launch<true, false>(tb, td, P, grid, s) : launch<true, true>(tb, td, P, grid, s);
  return opB == SPFY_OP_N ? launch<false, false>(tb, td, P, grid, s) : launch<false, true>(tb, td, P, grid, s);
}

}  // extern "C"

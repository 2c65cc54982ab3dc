// spmma.hxx -- sparsifyme::spmma: 2:4 magnitude prune + compress + structured-sparse GEMM.
//
// Same signature, defaults and return value as the reference
// (include/sparsify.me/spmma.hxx:21-118), but every cusparseLt call is replaced by our own
// sm_100a kernels behind the C ABI:
//   cusparseLtSpMMAPrune + PruneCheck + CompressedSize + Compress (:85-104) -> spfy_prune24
//        (prunes A in place and writes the compressed operand).  The pruning algorithm is the one the
//        reference asks the library for, CUSPARSELT_PRUNE_SPMMA_TILE (:86): best 2:4 pattern per 4x4
//        tile, two kept per row AND per column -- bit-exact with cusparseLt 0.7.1 including ties
//        (tests/golden/tile_*.npz).  Compile with -DSPARSIFYME_PRUNE_STRIP for the per-row
//        top-2-of-4 variant (CUSPARSELT_PRUNE_SPMMA_STRIP, also bit-exact), which is ONE fused
//        bandwidth-bound pass instead of two.
//   cusparseLtMatmul (:106-114)                                             -> spfy_spmma
//        (tcgen05.mma.sp, TMA-fed, fp32 accumulation in TMEM)
// Operands are row-major: A m x k (ld k), B k x n (ld n), C m x n (ld n); D aliases C (:52-64).
// Deliberate differences from the reference's defects: the element type follows
// type_t (the reference hard-codes CUDA_R_16F, :40) -- float buffers are converted to fp16 on
// the device, multiplied, and converted back; accumulation is fp32 (reference: fp16, :41).
// `batch_size` is accepted and unused exactly like the reference (:29).
#pragma once
#include <cstddef>
#include <iostream>
#include <vector>

#include <sparsify.me/detail/cabi.hxx>
#include <sparsify.me/util/util.hxx>

namespace sparsifyme {

template <typename type_t>
std::vector<float> spmma(
    type_t* dA,
    type_t* dB,
    type_t* dC,
    std::size_t m,
    std::size_t n,
    std::size_t k,
    std::size_t batch_size,
    cusparseOperation_t transpose_a = CUSPARSE_OPERATION_NON_TRANSPOSE,
    cusparseOperation_t transpose_b = CUSPARSE_OPERATION_NON_TRANSPOSE,
    float alpha = 1.0f,
    float beta = 0.0f) {
  (void)batch_size;
  constexpr bool is_f32 = std::is_same<type_t, float>::value;
  static_assert(is_f32 || std::is_same<type_t, __half>::value || std::is_same<type_t, __nv_bfloat16>::value,
                "spmma: type_t must be __half, __nv_bfloat16 or float");
  const int dtype = is_f32 ? SPFY_F16 : detail::dtype_of<type_t>::value;
  cudaStream_t stream = nullptr;
  spfy_stream_t s = reinterpret_cast<spfy_stream_t>(stream);
  detail::lazy_init();  // device code + scratch pool ready before any timer starts (cf. spmma.hxx:51-80)
  util::timer_t t;

  if (m % 8 != 0 || n % 8 != 0 || k % 8 != 0)  // same message policy as the reference (:45-49)
    std::cerr << "Invalid matrix sizes for data type __half. Rows and columns must be divisible by 8."
              << std::endl;
  // The reference hands transpose_a to cusparseLtMatmulDescriptorInit with A described as m x k (:56-69): for
  // op(A) = A^T the shapes only agree when m == k, and cusparseLt 0.7.1 has no transposed *structured* operand
  // for row-major A at all, so the reference's call chain fails and leaves C untouched.  Same here: nothing is
  // pruned or multiplied, the failure is reported, and the three timings are zero.
  if (transpose_a != CUSPARSE_OPERATION_NON_TRANSPOSE) {
    std::cerr << "sparsify.me: spmma: transpose_a is not supported (SPFY_E_UNSUPPORTED); nothing was computed."
              << std::endl;
#ifdef SPARSIFYME_STRICT
    throw std::runtime_error("spmma: transpose_a is not supported");
#endif
    return {0.f, 0.f, 0.f};
  }

  // float instantiation (the reference driver's: examples/spmma.cu:24): fp16 images of A, B, C
  const bool opb_t = transpose_b != CUSPARSE_OPERATION_NON_TRANSPOSE;
  detail::scratch hA(is_f32 ? m * k * 2 : 0, stream), hB(is_f32 ? k * n * 2 : 0, stream),
      hC(is_f32 ? m * n * 2 : 0, stream);
  void* A16 = dA;
  void* B16 = dB;
  void* C16 = dC;
  if (is_f32) {
    A16 = hA.ptr; B16 = hB.ptr; C16 = hC.ptr;
    detail::ok(spfy_convert(SPFY_F32, SPFY_F16, dA, A16, m * k, s), "spmma(convert A)");
    detail::ok(spfy_convert(SPFY_F32, SPFY_F16, dB, B16, k * n, s), "spmma(convert B)");
    if (beta != 0.f) detail::ok(spfy_convert(SPFY_F32, SPFY_F16, dC, C16, m * n, s), "spmma(convert C)");
  }

  std::size_t vals_bytes = 0, meta_bytes = 0;
  detail::ok(spfy_compressed_bytes(dtype, m, k, SPFY_LAYOUT_SM100, &vals_bytes, &meta_bytes), "spmma");
  detail::scratch vals(vals_bytes, stream), meta(meta_bytes, stream);

  // --- prune (+ compress, fused): A is pruned in place like cusparseLtSpMMAPrune(dA, dA) ---
  t.begin(stream);
#ifdef SPARSIFYME_PRUNE_STRIP
  constexpr int prune_alg = SPFY_PRUNE_STRIP_MAG;
#else
  constexpr int prune_alg = SPFY_PRUNE_TILE_MAG;
#endif
  detail::ok(spfy_prune24(dtype, prune_alg, SPFY_LAYOUT_SM100, A16, k, A16, k, vals.ptr, meta.ptr,
                          nullptr, m, k, s),
             "spmma(prune)");
  if (is_f32) detail::ok(spfy_convert(SPFY_F16, SPFY_F32, A16, dA, m * k, s), "spmma(convert A back)");
  const float prune_ms = t.end(stream);

  // --- compress: already produced by the fused kernel; the phase is kept for the 3-entry result ---
  t.begin(stream);
  const float compress_ms = t.end(stream);

  // --- multiply ---
  t.begin(stream);
  detail::ok(spfy_spmma(dtype, opb_t ? SPFY_OP_T : SPFY_OP_N, m, n, k, alpha, vals.ptr, meta.ptr, B16,
                        opb_t ? k : n, beta, C16, n, C16, n, nullptr, 0, s),
             "spmma(matmul)");
  if (is_f32) detail::ok(spfy_convert(SPFY_F16, SPFY_F32, C16, dC, m * n, s), "spmma(convert C back)");
  const float mul_ms = t.end(stream);

  return {prune_ms, compress_ms, mul_ms};
}

}  // namespace sparsifyme

// gemm.hxx -- sparsifyme::batched::gemm: the dense comparator the reference times next to its sparse
// paths (include/sparsify.me/gemm.hxx:25-195, caller examples/gemm.cu:93-95).
//
// Reference signature kept verbatim.  cublas{H,S}gemmBatched (:51, :133) is replaced by the dense tcgen05 GEMM
// of libsparsifyme_b200.so (spfy_gemm_batched): column-major, lda = m, ldb = k, ldc = m like the reference's
// call (:79-81); __half / __nv_bfloat16 run kind::f16 with fp32 accumulation, float runs 3xTF32 (fp32-level
// accuracy; -DSPARSIFYME_GEMM_FAST selects one TF32 product like cuBLAS's TF32 mode).  `cublasOperation_t` is
// only a parameter TYPE here.  double has no tensor-core path on this part: that instantiation alone still
// forwards to cublasDgemmBatched (:186) and is the only one that needs -lcublas.
//
// A_ptrs / B_ptrs / C_ptrs are DEVICE arrays of device pointers (examples/gemm.cu:64-88).  Tensor maps are
// built on the host, so the arrays are copied back once -- before the timer starts, next to the library
// warm-up, where the reference creates its cuBLAS handle (:48-49).
#pragma once
#include <cublas_v2.h>  // cublasOperation_t (types; cublasDgemmBatched for the double instantiation only)
#include <cuda_fp16.h>

#include <cstddef>
#include <cstdio>
#include <type_traits>
#include <vector>

#include <sparsify.me/detail/cabi.hxx>
#include <sparsify.me/util/util.hxx>

namespace sparsifyme {
namespace batched {

template <typename type_t>
float gemm(type_t** A_ptrs,
           type_t** B_ptrs,
           type_t** C_ptrs,
           std::size_t m,
           std::size_t n,
           std::size_t k,
           std::size_t batch_size,
           cublasOperation_t transpose_a = CUBLAS_OP_N,
           cublasOperation_t transpose_b = CUBLAS_OP_N,
           type_t alpha = (type_t)1.0f,
           type_t beta = (type_t)0.0f) {
  if (batch_size == 0) return 0.f;
  if constexpr (std::is_same<type_t, double>::value) {
    cublasHandle_t handle;
    cublasCreate(&handle);
    util::timer_t timer;
    timer.begin();
    cublasStatus_t status = cublasDgemmBatched(handle, transpose_a, transpose_b, (int)m, (int)n, (int)k, &alpha, A_ptrs,
                                               (int)m, B_ptrs, (int)k, &beta, C_ptrs, (int)m, (int)batch_size);
    const float ms = timer.end();
    if (status != CUBLAS_STATUS_SUCCESS) std::printf("cublas error: %d\n", (int)status);
    cublasDestroy(handle);
    return ms;
  } else {
    static_assert(detail::dtype_of<type_t>::value >= 0, "batched::gemm: type_t must be float, double, __half or __nv_bfloat16");
    cudaStream_t stream = nullptr;
    detail::lazy_init();
    std::vector<const void*> h(3 * batch_size);
    detail::cuda_ok(cudaMemcpy(h.data(), A_ptrs, batch_size * sizeof(void*), cudaMemcpyDeviceToHost), "batched::gemm");
    detail::cuda_ok(cudaMemcpy(h.data() + batch_size, B_ptrs, batch_size * sizeof(void*), cudaMemcpyDeviceToHost),
                    "batched::gemm");
    detail::cuda_ok(cudaMemcpy(h.data() + 2 * batch_size, C_ptrs, batch_size * sizeof(void*), cudaMemcpyDeviceToHost),
                    "batched::gemm");
#ifdef SPARSIFYME_GEMM_FAST
    const int precision = SPFY_GEMM_FAST;
#else
    const int precision = SPFY_GEMM_PRECISE;
#endif
    // the reference hands cuBLAS lda = m, ldb = k whatever the transposes are (:79-81); an operand stored
    // transposed has the other extent as its leading dimension, which is what is passed here
    const std::size_t lda = transpose_a == CUBLAS_OP_N ? m : k, ldb = transpose_b == CUBLAS_OP_N ? k : n;
    const int op_a = transpose_a == CUBLAS_OP_N ? SPFY_OP_N : SPFY_OP_T, op_b = transpose_b == CUBLAS_OP_N ? SPFY_OP_N : SPFY_OP_T;
    std::size_t ws_bytes = 0;
    detail::ok(spfy_gemm_workspace_bytes(detail::dtype_of<type_t>::value, op_a, op_b, m, n, k, lda, ldb, batch_size, &ws_bytes),
               "batched::gemm");
    detail::scratch ws(ws_bytes, stream);
    util::timer_t timer;
    timer.begin(stream);
    detail::ok(spfy_gemm_batched(detail::dtype_of<type_t>::value, precision, op_a, op_b, m, n, k, (float)alpha, h.data(), lda,
                                 h.data() + batch_size, ldb, (float)beta,
                                 const_cast<void* const*>(reinterpret_cast<const void* const*>(h.data() + 2 * batch_size)), m,
                                 batch_size, ws.ptr, ws_bytes, reinterpret_cast<spfy_stream_t>(stream)),
               "batched::gemm");
    return timer.end(stream);
  }
}

}  // namespace batched
}  // namespace sparsifyme

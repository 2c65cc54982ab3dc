// gemm.hxx -- sparsifyme::batched::gemm: the dense comparator the reference times next to its
// sparse paths (include/sparsify.me/gemm.hxx:25-195).  OUT OF SCOPE for the hot path (SURVEY.md
// C4): it stays a plain cuBLAS pointer-array batched GEMM so that examples/gemm.cu keeps
// building; column-major, lda = m, ldb = k, ldc = m (:79-81).
#pragma once
#include <cublas_v2.h>
#include <cuda_fp16.h>

#include <cstddef>
#include <cstdio>

#include <sparsify.me/util/util.hxx>

namespace sparsifyme {
namespace batched {
namespace detail_gemm {
inline cublasStatus_t run(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                          const __half* alpha, __half** A, int lda, __half** B, int ldb, const __half* beta,
                          __half** C, int ldc, int nb) {
  return cublasHgemmBatched(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb);
}
inline cublasStatus_t run(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                          const float* alpha, float** A, int lda, float** B, int ldb, const float* beta,
                          float** C, int ldc, int nb) {
  return cublasSgemmBatched(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb);
}
inline cublasStatus_t run(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                          const double* alpha, double** A, int lda, double** B, int ldb, const double* beta,
                          double** C, int ldc, int nb) {
  return cublasDgemmBatched(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb);
}
}  // namespace detail_gemm

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
  cublasHandle_t handle;
  cublasCreate(&handle);
  util::timer_t timer;
  timer.begin();
  cublasStatus_t status = detail_gemm::run(handle, transpose_a, transpose_b, (int)m, (int)n, (int)k, &alpha, A_ptrs,
                                           (int)m, B_ptrs, (int)k, &beta, C_ptrs, (int)m, (int)batch_size);
  const float ms = timer.end();
  if (status != CUBLAS_STATUS_SUCCESS) std::printf("cublas error: %d\n", (int)status);
  cublasDestroy(handle);
  return ms;
}

}  // namespace batched
}  // namespace sparsifyme

// examples/compare.cu -- bin/compare shapes.csv > compare.csv : the table the reference assembles with one
// subprocess per operator and layer (examples/profiling.py:20-44 -> examples/compare.csv:1,
// "layer,m,n,k,b,gemm,prune,spmm"), produced in ONE process through the public header API, and extended by
// the two operators of the north star that the reference's sweep never reached (profiling/spmm_timing.cu:64-66
// does not compile at HEAD): the 2:4 spmma phases and the batched COO SpMM.
//
// Operand shapes follow the reference drivers exactly (the CSV row (m, n, k, b) is passed straight through,
// examples/profiling.py:39-41): gemm and spmm multiply an m x k A by a k x n B per batch element in fp32;
// prune zeroes an m x k fp32 matrix positionally; spmma prunes + multiplies an fp16 m x k A (k padded to 8);
// coo keeps ~10 % of one shared m x k A (the 90 % sparsity of profiling/python/gemm_coo_compare.py:7).
// Every operator is called once un-timed (module load, plan caches) and once timed; times are milliseconds.
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/containers/ell.hxx>
#include <sparsify.me/gemm.hxx>
#include <sparsify.me/sparsify.hxx>
#include <sparsify.me/spmm.hxx>
#include <sparsify.me/spmma.hxx>
#include <sparsify.me/util/gen.hxx>
#include <sparsify.me/util/util.hxx>

using namespace sparsifyme;

static float time_gemm(std::size_t m, std::size_t n, std::size_t k, std::size_t b) {
  thrust::device_vector<float> A(b * m * k), B(b * k * n), C(b * m * n, 0.f);
  util::random::uniform_distribution(A, 0.f, 1.f);
  util::random::uniform_distribution(B, 0.f, 1.f);
  thrust::host_vector<float*> hA(b), hB(b), hC(b);
  for (std::size_t i = 0; i < b; ++i) {
    hA[i] = A.data().get() + i * m * k;
    hB[i] = B.data().get() + i * k * n;
    hC[i] = C.data().get() + i * m * n;
  }
  thrust::device_vector<float*> dA = hA, dB = hB, dC = hC;
  batched::gemm(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, b);
  return batched::gemm(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, b);
}

static float time_prune(std::size_t m, std::size_t k) {
  thrust::device_vector<float> w(m * k);
  util::random::uniform_distribution(w, 0.f, 1.f);
  thrust::device_vector<std::size_t> mask(m * k);
  sparsify<2, 2>(w.data().get(), mask.data().get(), m, k);
  util::timer_t t;
  t.begin();
  sparsify<2, 2>(w.data().get(), mask.data().get(), m, k);
  return t.end();
}

static float time_spmm_bell(std::size_t m, std::size_t n, std::size_t k, std::size_t b) {
  const std::size_t block = 2;
  std::vector<ell_t<float, memory_space_t::device>> d_As(b);
  ell_t<float, memory_space_t::host> h_A;  // one host image, copied b times (the driver draws b of them)
  h_A.rows = m;
  h_A.cols = k;
  h_A.block_size = block;
  h_A.ell_cols = std::max<std::size_t>(block, (k / 2) / block * block);
  h_A.blocked_rows = util::ceil_div(h_A.rows, h_A.block_size);
  h_A.blocked_cols = util::ceil_div(h_A.ell_cols, h_A.block_size);
  h_A.num_blocks = h_A.blocked_rows * h_A.blocked_cols;
  h_A.values.resize(h_A.rows * h_A.ell_cols);
  for (std::size_t i = 0; i < h_A.values.size(); ++i) h_A.values[i] = static_cast<float>(i % 7) - 3.f;
  h_A.column_indices.resize(h_A.num_blocks);
  const std::size_t pool = util::ceil_div(k, block);
  for (std::size_t br = 0; br < h_A.blocked_rows; ++br)
    for (std::size_t bc = 0; bc < h_A.blocked_cols; ++bc)  // ascending, distinct: every second block column, rotated
      h_A.column_indices[br * h_A.blocked_cols + bc] = (2 * bc + (br & 1)) % pool;
  for (std::size_t i = 0; i < b; ++i) d_As[i] = h_A;
  thrust::device_vector<float> d_B(k * n);
  util::random::uniform_distribution(d_B, 0.f, 1.f);
  std::vector<thrust::device_vector<float>> d_Cs(b);
  std::vector<float*> C_ptrs(b);
  for (std::size_t i = 0; i < b; ++i) {
    d_Cs[i].resize(m * n);
    C_ptrs[i] = d_Cs[i].data().get();
  }
  batched::spmm(d_As.data(), d_B.data().get(), C_ptrs.data(), m, n, k, b);
  return batched::spmm(d_As.data(), d_B.data().get(), C_ptrs.data(), m, n, k, b);
}

static std::vector<float> time_spmma(std::size_t m, std::size_t n, std::size_t k, std::size_t b) {
  const std::size_t K = (k + 7) / 8 * 8, N = (n + 7) / 8 * 8;
  thrust::device_vector<float> fa(m * K), fb(K * N);
  util::random::uniform_distribution(fa, -1.f, 1.f);
  util::random::uniform_distribution(fb, -1.f, 1.f);
  thrust::device_vector<__half> A(m * K), B(K * N), C(m * N);
  spfy_convert(SPFY_F32, SPFY_F16, fa.data().get(), A.data().get(), m * K, nullptr);
  spfy_convert(SPFY_F32, SPFY_F16, fb.data().get(), B.data().get(), K * N, nullptr);
  spmma(A.data().get(), B.data().get(), C.data().get(), m, N, K, b);
  return spmma(A.data().get(), B.data().get(), C.data().get(), m, N, K, b);
}

static float time_coo(std::size_t m, std::size_t n, std::size_t k, std::size_t b) {
  // one shared A with every 10th entry of a row kept (sorted by row, column), B_b k x n, C_b m x n column-major
  std::vector<int> rows, cols;
  std::vector<float> vals;
  for (std::size_t i = 0; i < m; ++i)
    for (std::size_t j = i % 10; j < k; j += 10) {
      rows.push_back((int)i);
      cols.push_back((int)j);
      vals.push_back(static_cast<float>((i + j) % 5) - 2.f);
    }
  thrust::device_vector<int> d_rows = rows, d_cols = cols;
  thrust::device_vector<float> d_vals = vals, d_B(b * k * n), d_C(b * m * n, 0.f);
  util::random::uniform_distribution(d_B, 0.f, 1.f);
  float* dC = d_C.data().get();
  batched::strided_coo<float>(m, k, vals.size(), k, n, b, d_rows.data().get(), d_cols.data().get(),
                              d_vals.data().get(), d_B.data().get(), &dC);
  return batched::strided_coo<float>(m, k, vals.size(), k, n, b, d_rows.data().get(), d_cols.data().get(),
                                     d_vals.data().get(), d_B.data().get(), &dC);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " shapes.csv" << std::endl;
    return EXIT_FAILURE;
  }
  std::vector<util::mat_sz> shapes;
  try {
    shapes = util::read_shapes(argv[1]);
  } catch (const char* msg) {
    std::cerr << msg << std::endl;
    return EXIT_FAILURE;
  }
  std::cout << "layer,m,n,k,b,gemm,prune,spmm,spmma_prune,spmma_compress,spmma_mul,coo90" << std::endl;
  std::size_t layer = 0;
  for (const auto& s : shapes) {
    const std::size_t m = std::get<0>(s), n = std::get<1>(s), k = std::get<2>(s), b = std::get<3>(s);
    const float g = time_gemm(m, n, k, b);
    const float p = time_prune(m, k);
    const float e = time_spmm_bell(m, n, k, b);
    const std::vector<float> a = time_spmma(m, n, k, b);
    const float c = time_coo(m, n, k, b);
    std::cout << layer++ << "," << m << "," << n << "," << k << "," << b << "," << g << "," << p << "," << e << ","
              << a[0] << "," << a[1] << "," << a[2] << "," << c << std::endl;
  }
  return 0;
}

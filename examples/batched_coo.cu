// examples/batched_coo.cu -- bin/batched_coo m n k b : one COO matrix A (m x k, ~50% random
// non-zeros) times b column-major B_i (k x n), prints the elapsed milliseconds (same CLI and
// stdout as the reference driver, examples/batched_coo.cu:31-112).  The reference driver sizes the
// index arrays by rows/cols instead of nnz and derives nnz from n (:46,:56,:64); here A is a
// proper row-sorted COO with nnz = ceil(m*k/2) distinct entries.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <random>
#include <vector>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/spmm.hxx>
#include <sparsify.me/util/gen.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using type_t = float;
  std::size_t m = 4, n = 4, k = 4, num_batches = 1;
  if (argc >= 5) {
    m = std::strtoull(argv[1], nullptr, 10);
    n = std::strtoull(argv[2], nullptr, 10);
    k = std::strtoull(argv[3], nullptr, 10);
    num_batches = std::strtoull(argv[4], nullptr, 10);
  }
  const std::size_t A_num_rows = m, A_num_cols = k, B_num_rows = k, B_num_cols = n;
  // every other position of the row-major m x k grid, offset by a random phase per row
  std::mt19937 rng(12345);
  thrust::host_vector<int> hA_rows, hA_cols;
  for (std::size_t r = 0; r < m; ++r) {
    const std::size_t phase = rng() & 1;
    for (std::size_t c = phase; c < k; c += 2) {
      hA_rows.push_back((int)r);
      hA_cols.push_back((int)c);
    }
  }
  const std::size_t A_nnz = hA_rows.size();
  thrust::device_vector<int> dA_rows = hA_rows, dA_cols = hA_cols;
  thrust::device_vector<type_t> dA_values(A_nnz);
  sparsifyme::util::random::uniform_distribution(dA_values, -1.f, 1.f);
  thrust::device_vector<type_t> dB(num_batches * B_num_rows * B_num_cols);
  sparsifyme::util::random::uniform_distribution(dB, -1.f, 1.f);
  thrust::device_vector<type_t> dC(num_batches * A_num_rows * B_num_cols, 0.f);
  type_t* dC_ptr = dC.data().get();

  float elapsed = sparsifyme::batched::strided_coo(A_num_rows, A_num_cols, A_nnz, B_num_rows, B_num_cols, num_batches,
                                                   dA_rows.data().get(), dA_cols.data().get(), dA_values.data().get(),
                                                   dB.data().get(), &dC_ptr);
  std::cout << elapsed << std::endl;
  return 0;
}

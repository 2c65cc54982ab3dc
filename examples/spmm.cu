// examples/spmm.cu -- bin/spmm m n k b : batched blocked-ELL SpMM, prints the elapsed milliseconds
// (same CLI and stdout as the reference driver, examples/spmm.cu:24-118).  Construction of the
// operands follows the reference: block 2, ell_cols = k/2, values 1,2,3..., sorted unique random
// block-column ids per block row, B = 1,2,3... column-major, one C per batch element.
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <random>
#include <vector>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/containers/ell.hxx>
#include <sparsify.me/spmm.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using namespace sparsifyme;
  using type_t = float;
  std::size_t m = 4, n = 4, k = 4, batch_size = 1;
  if (argc >= 5) {
    m = std::strtoull(argv[1], nullptr, 10);
    n = std::strtoull(argv[2], nullptr, 10);
    k = std::strtoull(argv[3], nullptr, 10);
    batch_size = std::strtoull(argv[4], nullptr, 10);
  }
  const std::size_t block = 2;
  std::mt19937 rng(std::random_device{}());
  std::vector<ell_t<type_t, memory_space_t::device>> d_As(batch_size);
  for (std::size_t b = 0; b < batch_size; ++b) {
    ell_t<type_t, memory_space_t::host> h_A;
    h_A.rows = m;
    h_A.cols = k;
    h_A.block_size = block;
    h_A.ell_cols = std::max<std::size_t>(block, (k / 2) / block * block);
    h_A.blocked_rows = util::ceil_div(h_A.rows, h_A.block_size);
    h_A.blocked_cols = util::ceil_div(h_A.ell_cols, h_A.block_size);
    h_A.num_blocks = h_A.blocked_rows * h_A.blocked_cols;
    h_A.values.resize(h_A.rows * h_A.ell_cols);
    for (std::size_t i = 0; i < h_A.values.size(); ++i) h_A.values[i] = static_cast<type_t>(i + 1);
    h_A.column_indices.resize(h_A.num_blocks);
    std::vector<std::size_t> pool(util::ceil_div(k, block));
    for (std::size_t br = 0; br < h_A.blocked_rows; ++br) {
      std::iota(pool.begin(), pool.end(), 0);
      std::shuffle(pool.begin(), pool.end(), rng);
      std::sort(pool.begin(), pool.begin() + h_A.blocked_cols);
      for (std::size_t bc = 0; bc < h_A.blocked_cols; ++bc) h_A.column_indices[br * h_A.blocked_cols + bc] = pool[bc];
    }
    d_As[b] = h_A;
  }
  thrust::host_vector<type_t> h_B(k * n);
  for (std::size_t i = 0; i < h_B.size(); ++i) h_B[i] = static_cast<type_t>(i + 1);
  thrust::device_vector<type_t> d_B = h_B;
  std::vector<thrust::device_vector<type_t>> d_Cs(batch_size);
  std::vector<type_t*> C_ptrs(batch_size);
  for (std::size_t b = 0; b < batch_size; ++b) {
    d_Cs[b].resize(m * n);
    C_ptrs[b] = d_Cs[b].data().get();
  }
  float elapsed = batched::spmm(d_As.data(), d_B.data().get(), C_ptrs.data(), m, n, k, batch_size);
  std::cout << elapsed << std::endl;
  return 0;
}

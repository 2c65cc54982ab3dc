// containers/ell.hxx -- blocked-ELL matrix, the A operand of sparsifyme::batched::spmm.
// Field names and meaning follow the reference container
// (include/sparsify.me/containers/ell.hxx:24-33) because the drivers fill them directly
// (examples/spmm.cu:45-84): rows x cols matrix, square blocks of `block_size`, `ell_cols`
// stored columns per row; column_indices holds one block-column id per stored block
// ([blocked_rows x blocked_cols], std::size_t), values is [rows x ell_cols] row-major.
#pragma once
#include <cstddef>
#include <iostream>

#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/containers/memory.hxx>
#include <sparsify.me/containers/vector.hxx>

namespace sparsifyme {
template <typename type_t = float, memory_space_t space = memory_space_t::device>
struct ell_t {
  std::size_t rows = 0, cols = 0, block_size = 0;
  std::size_t ell_cols = 0;
  std::size_t blocked_rows = 0;  // rows / block_size
  std::size_t blocked_cols = 0;  // ell_cols / block_size
  std::size_t num_blocks = 0;    // blocked_rows * blocked_cols

  vector_t<std::size_t, space> column_indices;
  vector_t<type_t, space> values;

  ell_t() = default;

  // cross-space assignment (host -> device upload and back), one thrust copy per array
  template <memory_space_t other>
  ell_t& operator=(const ell_t<type_t, other>& src) {
    rows = src.rows;
    cols = src.cols;
    block_size = src.block_size;
    ell_cols = src.ell_cols;
    blocked_rows = src.blocked_rows;
    blocked_cols = src.blocked_cols;
    num_blocks = src.num_blocks;
    column_indices = src.column_indices;
    values = src.values;
    return *this;
  }

  void print() const {
    thrust::host_vector<std::size_t> idx = column_indices;
    thrust::host_vector<type_t> val = values;
    std::cout << "A-Matrix\n\t(rows, cols) = " << rows << ", " << cols << "\n\tELL columns = " << ell_cols
              << "\n\tBlock Size = " << block_size << "\n\tNumber of Blocks = " << num_blocks
              << "\n\tColumn Idx = ";
    for (std::size_t i = 0; i < idx.size(); ++i) std::cout << idx[i] << " ";
    std::cout << "\n\tValues = ";
    for (std::size_t i = 0; i < val.size(); ++i) std::cout << static_cast<float>(val[i]) << " ";
    std::cout << std::endl;
  }
};
}  // namespace sparsifyme

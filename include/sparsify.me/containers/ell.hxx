// containers/ell.hxx -- blocked-ELL matrix, the A operand of sparsifyme::batched::spmm.
//
// The drivers fill the container member by member (examples/spmm.cu:45-84 of the reference), so the
// member names and their meaning are the interface: a rows x cols matrix cut into square blocks of
// `block_size`; every row stores `ell_cols` values; `column_indices` holds one block-column id per stored
// block, [blocked_rows x blocked_cols] as std::size_t (reference: containers/ell.hxx:24-33); `values` is
// [rows x ell_cols] row-major.  A negative id (as int64) marks a padding block.
//
// The geometry lives in a plain base so that moving a matrix between memory spaces is "copy the shape,
// then assign the two arrays" -- vector_t's cross-space assignment does the transfer.
#pragma once
#include <cstddef>
#include <ostream>
#include <iostream>

#include <thrust/host_vector.h>

#include <sparsify.me/containers/memory.hxx>
#include <sparsify.me/containers/vector.hxx>

namespace sparsifyme {

struct ell_shape_t {
  std::size_t rows = 0;
  std::size_t cols = 0;
  std::size_t block_size = 0;
  std::size_t ell_cols = 0;      // stored columns per row
  std::size_t blocked_rows = 0;  // ceil(rows / block_size)
  std::size_t blocked_cols = 0;  // ell_cols / block_size
  std::size_t num_blocks = 0;    // blocked_rows * blocked_cols
};

template <typename type_t = float, memory_space_t space = memory_space_t::device>
struct ell_t : ell_shape_t {
  vector_t<std::size_t, space> column_indices;
  vector_t<type_t, space> values;

  ell_t() = default;

  template <memory_space_t from>
  ell_t(const ell_t<type_t, from>& other) { *this = other; }

  // host <-> device: the shape is copied by value, the arrays by thrust
  template <memory_space_t from>
  ell_t& operator=(const ell_t<type_t, from>& other) {
    static_cast<ell_shape_t&>(*this) = static_cast<const ell_shape_t&>(other);
    column_indices = other.column_indices;
    values = other.values;
    return *this;
  }

  void print(std::ostream& os = std::cout) const {
    const thrust::host_vector<std::size_t> ids = column_indices;
    const thrust::host_vector<type_t> vals = values;
    os << "A-Matrix\n"
       << "\t(rows, cols) = " << rows << ", " << cols << '\n'
       << "\tELL columns = " << ell_cols << '\n'
       << "\tBlock Size = " << block_size << '\n'
       << "\tNumber of Blocks = " << num_blocks << '\n';
    os << "\tColumn Idx = ";
    for (const std::size_t id : ids) os << id << ' ';
    os << "\n\tValues = ";
    for (const type_t& v : vals) os << static_cast<float>(v) << ' ';
    os << std::endl;
  }
};

}  // namespace sparsifyme

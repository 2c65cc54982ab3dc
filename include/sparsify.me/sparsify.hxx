// sparsify.hxx -- sparsifyme::sparsify<BLK_M, BLK_N> and the magnitude-pruning entry points.
//
// `sparsify` keeps the reference signature and its exact positional semantics
// (reference: include/sparsify.me/sparsify.hxx:24-82): mask <- 1, then in every linear block of
// BLK_M*BLK_N elements zero floor(BLK_M*BLK_N*sparsity_factor) offsets of the sequence
// h + w*BLK_N (h outer, w inner) in `weights` and `mask`.  It is one fused sm_100a kernel behind
// spfy_prune_blocks_ref instead of thrust::fill_n + thrust::transform.
//
// The reference marks the selection rule as a <todo> (:58-59); the value-aware rule it hands to
// cusparseLt elsewhere (spmma.hxx:86) is exposed here as `prune24`.
#pragma once
#include <cstddef>

#include <sparsify.me/detail/cabi.hxx>
#include <sparsify.me/util/util.hxx>

namespace sparsifyme {

template <std::size_t BLK_M = 2, std::size_t BLK_N = 2, typename type_t>
void sparsify(type_t* weights,
              std::size_t* mask,
              std::size_t const& m,
              std::size_t const& n,
              float sparsity_factor = 0.5,
              cudaStream_t stream = 0) {
  static_assert(detail::dtype_of<type_t>::value >= 0, "sparsify: unsupported element type");
  static_assert(sizeof(std::size_t) == sizeof(std::uint64_t), "mask words are 64-bit");
  detail::lazy_init();  // first use on this device loads the device code (inside the caller's timer, like the reference)
  detail::ok(spfy_prune_blocks_ref(detail::dtype_of<type_t>::value, weights,
                                   reinterpret_cast<std::uint64_t*>(mask), m, n, BLK_M, BLK_N,
                                   sparsity_factor, reinterpret_cast<spfy_stream_t>(stream)),
             "sparsify");
  // thrust::cuda::par.on(stream) in the reference returns only when the work is done
  detail::cuda_ok(cudaStreamSynchronize(stream), "sparsify");
}

// 2:4 magnitude prune of a rows x cols row-major matrix, in place: per group of 4 consecutive
// columns keep the 2 largest |x| (tie -> lower index).  Optionally fills the reference-style
// std::size_t keep mask (pass nullptr to skip it).  type_t is __half or __nv_bfloat16.
template <typename type_t>
void prune24(type_t* weights, std::size_t* mask, std::size_t rows, std::size_t cols,
             cudaStream_t stream = 0) {
  detail::ok(spfy_prune24(detail::dtype_of<type_t>::value, SPFY_PRUNE_STRIP_MAG, SPFY_LAYOUT_CANONICAL,
                          weights, cols, weights, cols, nullptr, nullptr,
                          reinterpret_cast<std::uint64_t*>(mask), rows, cols,
                          reinterpret_cast<spfy_stream_t>(stream)),
             "prune24");
}

}  // namespace sparsifyme

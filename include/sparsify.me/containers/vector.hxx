// containers/vector.hxx -- vector_t<T, space>: thrust::host_vector or thrust::device_vector
// selected by memory space (reference: include/sparsify.me/containers/vector.hxx:18-23).
#pragma once
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <type_traits>

#include <sparsify.me/containers/memory.hxx>

namespace sparsifyme {
namespace detail {
template <typename T, memory_space_t S> struct vector_for { using type = thrust::device_vector<T>; };
template <typename T> struct vector_for<T, memory_space_t::host> { using type = thrust::host_vector<T>; };
}  // namespace detail

template <typename type_t, memory_space_t space>
using vector_t = typename detail::vector_for<type_t, space>::type;
}  // namespace sparsifyme

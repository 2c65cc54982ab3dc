// util/gen.hxx -- fill a thrust vector with U[begin, end) values
// (reference: include/sparsify.me/util/gen.hxx:8-21).  Counter-based: element i is a pure
// function of (i, begin, end), so device and host vectors get identical contents.
#pragma once
#include <cstdint>

#include <thrust/iterator/counting_iterator.h>
#include <thrust/transform.h>

namespace sparsifyme {
namespace util {
namespace random {

struct counter_uniform {
  double lo, span;
  __host__ __device__ double operator()(unsigned long long i) const {
    // splitmix64 finaliser -> 53 random bits -> [0, 1)
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return lo + span * (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
};

template <typename vector_t, typename type_t = typename vector_t::value_type>
void uniform_distribution(vector_t& input, type_t begin = 0.0f, type_t end = 1.0f) {
  const counter_uniform f{(double)begin, (double)end - (double)begin};
  thrust::transform(thrust::make_counting_iterator<unsigned long long>(0),
                    thrust::make_counting_iterator<unsigned long long>(input.size()), input.begin(),
                    [f] __host__ __device__(unsigned long long i) -> type_t { return (type_t)f(i); });
}

}  // namespace random
}  // namespace util
}  // namespace sparsifyme

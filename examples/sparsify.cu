// examples/sparsify.cu -- bin/sparsify m n : positional block pruning of an m x n fp32 matrix,
// prints the elapsed milliseconds (same CLI and stdout as the reference driver,
// reference: examples/sparsify.cu:19-54, parsed by examples/profiling.py:11-13).
#include <cstdlib>
#include <iostream>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/sparsify.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using type_t = float;
  std::size_t m = 8, n = 8;
  if (argc >= 3) {
    m = std::strtoull(argv[1], nullptr, 10);
    n = std::strtoull(argv[2], nullptr, 10);
  }
  thrust::host_vector<type_t> h_weights(m * n);
  for (std::size_t i = 0; i < h_weights.size(); ++i) h_weights[i] = sparsifyme::util::get_random<type_t>();
  thrust::device_vector<type_t> d_weights = h_weights;
  thrust::device_vector<std::size_t> d_mask(m * n);

  sparsifyme::util::timer_t timer;
  timer.begin();
  sparsifyme::sparsify<2, 2>(d_weights.data().get(), d_mask.data().get(), m, n);
  timer.end();
  std::cout << timer.milliseconds() << std::endl;
  return 0;
}

// examples/sweep.cu -- bin/sweep shapes.csv [ref|weights] : the sweep the reference intended in
// profiling/gemm_timing.cu and profiling/spmm_timing.cu (neither compiles at HEAD):
// for every row of a datasets/*.csv table run prune+compress and the 2:4 GEMM through the public
// header API and print CSV  m,n,k,b,prune,compress,spmma  (ms), the schema of
// profiling/gemm_timing.cu:41,110 extended by the spmma phases (spmm_timing.cu:64-66).
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>

#include <sparsify.me/spmma.hxx>
#include <sparsify.me/util/gen.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using namespace sparsifyme;
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " shapes.csv [ref|weights]" << std::endl;
    return EXIT_FAILURE;
  }
  const bool weights = argc < 3 || std::string(argv[2]) == "weights";
  std::vector<util::mat_sz> shapes;
  try {
    shapes = util::read_shapes(argv[1]);
  } catch (const char* msg) {
    std::cerr << msg << std::endl;
    return EXIT_FAILURE;
  }
  std::cout << "m,n,k,b,prune,compress,spmma" << std::endl;
  for (const auto& s : shapes) {
    const std::size_t m = std::get<0>(s), n = std::get<1>(s), k = std::get<2>(s), b = std::get<3>(s);
    // weights orientation: A = C_out x K weights, N = spatial * batch; ref: A = m x k as the drivers pass it
    const std::size_t M = weights ? n : m, N = weights ? m * b : n, K = (k + 7) / 8 * 8;
    thrust::device_vector<float> fa(M * K), fb(K * N);
    util::random::uniform_distribution(fa, -1.f, 1.f);
    util::random::uniform_distribution(fb, -1.f, 1.f);
    thrust::device_vector<__half> A(M * K), B(K * N), C(M * N);
    spfy_convert(SPFY_F32, SPFY_F16, fa.data().get(), A.data().get(), M * K, nullptr);
    spfy_convert(SPFY_F32, SPFY_F16, fb.data().get(), B.data().get(), K * N, nullptr);
    std::vector<float> t = spmma(A.data().get(), B.data().get(), C.data().get(), M, N, K, b);
    std::cout << m << "," << n << "," << k << "," << b << "," << t[0] << "," << t[1] << "," << t[2] << std::endl;
  }
  return 0;
}

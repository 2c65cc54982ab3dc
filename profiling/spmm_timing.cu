// profiling/spmm_timing.cu -- the 2:4 sparse-GEMM timing driver the reference sketches at profiling/spmm_timing.cu:15-66
// (it includes a header that does not exist, <sparsify.me/ampere.hxx>, and does not compile): same command line and
// output,
//
//     spmm_timing m n k b        ->  "<prune ms>, <compress ms>, <multiply ms>"
//
// through sparsifyme::spmma, which is what `ampere_spmm` became (include/sparsify.me/spmma.hxx:21-118).  Inputs are
// small integers drawn with std::rand like the reference's (:45-50), held in fp16 (the type the reference's
// descriptors declare, spmma.hxx:40); the compute-capability gate of :24-31 is gone -- the path needs sm_100.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/spmma.hxx>

int main(int argc, char** argv) {
  using type_t = __half;
  if (argc != 5) {
    std::cerr << "Invalid # of args. Usage: ./spmm_timing m n k b" << std::endl;
    return EXIT_FAILURE;
  }
  const std::size_t m = std::stoul(argv[1]), n = std::stoul(argv[2]), k = std::stoul(argv[3]), b = std::stoul(argv[4]);
  thrust::host_vector<type_t> hA(m * k), hB(k * n);
  for (auto& a : hA) a = type_t(static_cast<float>(std::rand() % 100));
  for (auto& v : hB) v = type_t(static_cast<float>(std::rand() % 100) / 128.f);
  thrust::device_vector<type_t> dA = hA, dB = hB, dC(m * n);
  const auto times = sparsifyme::spmma(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, b);
  std::cout << times[0] << ", " << times[1] << ", " << times[2] << std::endl;
  return EXIT_SUCCESS;
}

// examples/spmma.cu -- bin/spmma m n k b : 2:4 prune + compress + sparse GEMM, prints the three
// phase timings (same CLI and stdout lines as the reference driver, examples/spmma.cu:22-66).
// Differences: runs on compute capability 10.x (the reference insists on 8.0, :35-40) and uses
// fp16 data (the reference feeds float buffers to an fp16 descriptor, spmma.hxx:40).
#include <cstdlib>
#include <iostream>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/spmma.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using namespace sparsifyme;
  using type_t = __half;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, 0);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, 0);
  if (major != 10) {
    std::cerr << "\nthis build of spmma runs tcgen05.mma.sp and needs compute capability 10.x (found " << major << "."
              << minor << ")\n" << std::endl;
    return EXIT_FAILURE;
  }
  std::size_t m = 32, n = 32, k = 32, batch_size = 1;
  if (argc >= 5) {
    m = std::strtoull(argv[1], nullptr, 10);
    n = std::strtoull(argv[2], nullptr, 10);
    k = std::strtoull(argv[3], nullptr, 10);
    batch_size = std::strtoull(argv[4], nullptr, 10);
  }
  // like the reference (:48-59) the buffers hold `batch_size` problems; spmma multiplies the first
  thrust::host_vector<type_t> hA(m * k * batch_size), hB(k * n * batch_size);
  for (auto& a : hA) a = __float2half(util::get_random());
  for (auto& b : hB) b = __float2half(util::get_random());
  thrust::device_vector<type_t> dA = hA, dB = hB, dC(m * n * batch_size);

  std::vector<float> times = spmma(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, batch_size);
  std::cout << "Prune time: " << times[0] << " ms" << std::endl;
  std::cout << "Compress time: " << times[1] << " ms" << std::endl;
  std::cout << "Multiplication time: " << times[2] << " ms" << std::endl;
  return 0;
}

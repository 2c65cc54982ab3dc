// examples/spmma.cu -- bin/spmma m n k b : 2:4 prune + compress + sparse GEMM of one problem, printing the
// three phase times exactly as examples/profiling.py:8-17 of the reference parses them
// ("Prune time: … ms" / "Compress time: … ms" / "Multiplication time: … ms").
// Unlike the reference driver (examples/spmma.cu:35-40 insists on compute capability 8.0 and feeds float
// buffers to an fp16 descriptor, spmma.hxx:40) this one runs on 10.x and on real fp16 data.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/spmma.hxx>
#include <sparsify.me/util/util.hxx>

namespace {

bool device_is_blackwell() {
  int cc[2] = {0, 0};
  cudaDeviceGetAttribute(&cc[0], cudaDevAttrComputeCapabilityMajor, 0);
  cudaDeviceGetAttribute(&cc[1], cudaDevAttrComputeCapabilityMinor, 0);
  if (cc[0] == 10) return true;
  std::cerr << "\nthis build of spmma issues tcgen05.mma.sp and needs compute capability 10.x (found " << cc[0] << "."
            << cc[1] << ")\n" << std::endl;
  return false;
}

thrust::device_vector<__half> random_halves(std::size_t count) {
  thrust::host_vector<__half> h(count);
  for (std::size_t i = 0; i < count; ++i) h[i] = __float2half(sparsifyme::util::get_random());
  return h;
}

}  // namespace

int main(int argc, char** argv) {
  if (!device_is_blackwell()) return EXIT_FAILURE;
  std::size_t dims[4] = {32, 32, 32, 1};  // m n k b
  if (argc >= 5)
    for (int i = 0; i < 4; ++i) dims[i] = std::strtoull(argv[i + 1], nullptr, 10);
  const std::size_t m = dims[0], n = dims[1], k = dims[2], batch = dims[3];

  // the buffers are sized for `batch` problems like the reference's (:48-59); spmma multiplies the first
  thrust::device_vector<__half> A = random_halves(m * k * batch), B = random_halves(k * n * batch);
  thrust::device_vector<__half> C(m * n * batch);
  const std::vector<float> ms = sparsifyme::spmma(A.data().get(), B.data().get(), C.data().get(), m, n, k, batch);

  const char* phase[3] = {"Prune", "Compress", "Multiplication"};
  for (int i = 0; i < 3; ++i) std::cout << phase[i] << " time: " << ms[i] << " ms" << std::endl;
  return 0;
}

// examples/gemm.cu -- bin/gemm m n k b : dense batched GEMM comparator (cuBLAS), prints the elapsed
// milliseconds (same CLI and stdout as the reference driver, examples/gemm.cu:21-97).
#include <cstdlib>
#include <iostream>
#include <vector>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/gemm.hxx>
#include <sparsify.me/util/gen.hxx>
#include <sparsify.me/util/util.hxx>

int main(int argc, char** argv) {
  using type_t = float;
  std::size_t m = 4, n = 4, k = 4, batch_size = 1;
  if (argc >= 5) {
    m = std::strtoull(argv[1], nullptr, 10);
    n = std::strtoull(argv[2], nullptr, 10);
    k = std::strtoull(argv[3], nullptr, 10);
    batch_size = std::strtoull(argv[4], nullptr, 10);
  }
  thrust::device_vector<type_t> A(batch_size * m * k), B(batch_size * k * n), C(batch_size * m * n, 0.f);
  sparsifyme::util::random::uniform_distribution(A, 0.f, 1.f);
  sparsifyme::util::random::uniform_distribution(B, 0.f, 1.f);
  thrust::host_vector<type_t*> hA(batch_size), hB(batch_size), hC(batch_size);
  for (std::size_t b = 0; b < batch_size; ++b) {
    hA[b] = A.data().get() + b * m * k;
    hB[b] = B.data().get() + b * k * n;
    hC[b] = C.data().get() + b * m * n;
  }
  thrust::device_vector<type_t*> dA = hA, dB = hB, dC = hC;
  float elapsed = sparsifyme::batched::gemm(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, batch_size);
  std::cout << elapsed << std::endl;
  return 0;
}

// profiling/gemm_timing.cu -- the dense-GEMM sweep the reference sketches at profiling/gemm_timing.cu:20-110 (it does
// not compile at HEAD: a switch over a string selects a type alias): same command line and same CSV,
//
//     gemm_timing <f|h|d> <out.csv> [shapes.csv]        ->  "m,n,k,b,elapsed" per row of the shape table
//
// through sparsifyme::batched::gemm on the reference's operand shapes (examples/gemm.cu:60-95: one m x k A per batch
// element, ONE k x n B shared by the batch, column-major).  The table defaults to ../datasets/shapes.csv like the
// reference (:38).  Every shape is run once un-timed and once timed.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/gemm.hxx>
#include <sparsify.me/util/gen.hxx>
#include <sparsify.me/util/util.hxx>

using namespace sparsifyme;

template <typename type_t>
static int sweep(const std::vector<util::mat_sz>& shapes, std::ofstream& of) {
  of << "m,n,k,b,elapsed\n";
  for (const auto& sz : shapes) {
    const std::size_t m = std::get<0>(sz), n = std::get<1>(sz), k = std::get<2>(sz), b = std::get<3>(sz);
    // random operands, generated on the device in fp32 and narrowed where needed
    thrust::device_vector<float> fa(b * m * k), fb(k * n);
    util::random::uniform_distribution(fa, 0.f, 1.f);
    util::random::uniform_distribution(fb, 0.f, 1.f);
    thrust::device_vector<type_t> A(fa.begin(), fa.end()), B(fb.begin(), fb.end()), C(b * m * n);
    fa.clear();
    fa.shrink_to_fit();
    thrust::host_vector<type_t*> hA(b), hB(b), hC(b);
    for (std::size_t i = 0; i < b; ++i) {
      hA[i] = A.data().get() + i * m * k;
      hB[i] = B.data().get();  // one B for the whole batch
      hC[i] = C.data().get() + i * m * n;
    }
    thrust::device_vector<type_t*> dA = hA, dB = hB, dC = hC;
    batched::gemm(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, b);
    const float elapsed = batched::gemm(dA.data().get(), dB.data().get(), dC.data().get(), m, n, k, b);
    of << m << "," << n << "," << k << "," << b << "," << elapsed << "\n";
  }
  return EXIT_SUCCESS;
}

int main(int argc, char** argv) {
  if (argc != 3 && argc != 4) {
    std::cerr << "Invalid # of input arguments. Usage: ./gemm_timing float_precision (f,h,d) filename.csv [shapes.csv]"
              << std::endl;
    return EXIT_FAILURE;
  }
  std::vector<util::mat_sz> shapes;
  try {
    shapes = util::read_shapes(argc == 4 ? argv[3] : "../datasets/shapes.csv");
  } catch (const char* msg) {
    std::cerr << msg << std::endl;
    return EXIT_FAILURE;
  }
  std::ofstream of(argv[2]);
  if (!of.is_open()) {
    std::cerr << "cannot write " << argv[2] << std::endl;
    return EXIT_FAILURE;
  }
  switch (argv[1][0]) {
    case 'd': return sweep<double>(shapes, of);
    case 'h': return sweep<__half>(shapes, of);
    default: return sweep<float>(shapes, of);
  }
}

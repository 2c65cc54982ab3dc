// ref_sparsify_dump.cu -- runs the REFERENCE's own sparsifyme::sparsify<BLK_M,BLK_N>
// (compiled from /root/reference/include/sparsify.me/sparsify.hxx, not copied) on a
// deterministic input and dumps weights + mask so the CPU oracle
// (orc_prune_blocks_ref) and our kernel (spfy_prune_blocks_ref) can be checked
// bit-for-bit against it.  TEST INFRASTRUCTURE ONLY; needs a GPU to run.
//
// usage: ref_sparsify_dump <m> <n> <blk: 22|24|42|44> <sparsity_factor> <out.bin>
// input : weights[i] = 1 + (i % 251) as float        (never zero, so zeros mark pruning)
// output: float weights[m*n] followed by uint64 mask[m*n]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <sparsify.me/sparsify.hxx>  // the reference header

int main(int argc, char** argv) {
  if (argc != 6) {
    std::fprintf(stderr, "usage: %s m n blk sparsity out.bin\n", argv[0]);
    return 2;
  }
  std::size_t m = std::atoll(argv[1]), n = std::atoll(argv[2]);
  int blk = std::atoi(argv[3]);
  float sf = (float)std::atof(argv[4]);
  std::vector<float> h(m * n);
  for (std::size_t i = 0; i < m * n; ++i) h[i] = 1.0f + (float)(i % 251);
  thrust::device_vector<float> d(h.begin(), h.end());
  thrust::device_vector<std::size_t> mask(m * n);
  switch (blk) {
    case 22: sparsifyme::sparsify<2, 2>(d.data().get(), mask.data().get(), m, n, sf); break;
    case 42: sparsifyme::sparsify<4, 2>(d.data().get(), mask.data().get(), m, n, sf); break;
    case 44: sparsifyme::sparsify<4, 4>(d.data().get(), mask.data().get(), m, n, sf); break;
    case 11: sparsifyme::sparsify<1, 1>(d.data().get(), mask.data().get(), m, n, sf); break;
    default: std::fprintf(stderr, "unsupported blk %d\n", blk); return 2;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return 3;
  thrust::host_vector<float> hw = d;
  thrust::host_vector<std::size_t> hm = mask;
  FILE* f = std::fopen(argv[5], "wb");
  if (!f) return 4;
  std::fwrite(hw.data(), sizeof(float), m * n, f);
  std::fwrite(hm.data(), sizeof(std::size_t), m * n, f);
  std::fclose(f);
  return 0;
}

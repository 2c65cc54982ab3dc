// cusparselt_ref.cu -- comparator + golden generator that issues THE SAME cusparseLt call
// sequence the reference makes (include/sparsify.me/spmma.hxx:51-114) against the on-box
// cusparseLt 0.7.x.  TEST / BENCH INFRASTRUCTURE ONLY; never linked into the product.
//
// The reference header itself cannot be compiled against 0.7.x (PlanInit / CompressedSize /
// Compress changed signature) and it hard-codes fp16 descriptors on float data, so
// this is OUR driver around the library, following the reference's sequence:
//   Init -> StructuredDescriptorInit(A m x k ld k ROW 50%) -> DenseDescriptorInit(B k x n ld n ROW)
//   -> DenseDescriptorInit(C m x n ld n ROW) -> MatmulDescriptorInit(N, N, A, B, C, C)
//   -> AlgSelectionInit(DEFAULT) -> PlanInit -> GetWorkspace
//   -> [t] SpMMAPrune(in place) + PruneCheck -> [t] CompressedSize + Compress -> [t] Matmul
//
//   cusparselt_ref golden <m> <k> <n> <tile|strip> <out.bin>
//        deterministic fp16 inputs (splitmix64 counter RNG, see gen()); dumps
//        header{m,k,n,alg,valid}, A_in[m*k], A_pruned[m*k], B[k*n], D[m*n] (all uint16 fp16 bits)
//   cusparselt_ref prunefile <m> <k> <tile|strip> <f16|bf16> <in.bin> <out.bin>
//        prune a 16-bit matrix read from a file (tie-break probes, tests/golden/make_tile_probe.py)
//   cusparselt_ref sweep <shapes.csv> <batch>
//        weights orientation (M = n_csv, K = k_csv, N = m_csv*batch); per layer: prune / compress /
//        matmul ms (median of 5 after warm-up, cold L2); last line: one JSON object with totals
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cusparseLt.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(3);                                                                    \
    }                                                                                  \
  } while (0)
#define CKS(x)                                                                         \
  do {                                                                                 \
    cusparseStatus_t s_ = (x);                                                         \
    if (s_ != CUSPARSE_STATUS_SUCCESS) {                                               \
      std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cusparseGetErrorString(s_)); \
      return 4;                                                                        \
    }                                                                                  \
  } while (0)

// element i of tensor t: U(-1,1) quantised to multiples of 1/64 (so that fp16 ties are frequent
// and every value is exact in fp16); same formula in tests/golden/make_golden.py
static inline float gen(uint64_t t, uint64_t i) {
  uint64_t z = (t * 0x632BE59BD9B4E019ull + i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  int q = (int)(z >> 57) - 64;  // [-64, 63]
  return (float)q / 64.0f;
}

struct Problem {
  cusparseLtHandle_t* h;
  cusparseLtMatDescriptor_t dA, dB, dC;
  cusparseLtMatmulDescriptor_t mm;
  cusparseLtMatmulAlgSelection_t alg;
  cusparseLtMatmulPlan_t plan;
  size_t ws = 0;
  int init(cusparseLtHandle_t* handle, int64_t m, int64_t n, int64_t k, cudaDataType dt = CUDA_R_16F) {
    h = handle;
    CKS(cusparseLtStructuredDescriptorInit(h, &dA, m, k, k, 16, dt, CUSPARSE_ORDER_ROW,
                                           CUSPARSELT_SPARSITY_50_PERCENT));
    CKS(cusparseLtDenseDescriptorInit(h, &dB, k, n, n, 16, dt, CUSPARSE_ORDER_ROW));
    CKS(cusparseLtDenseDescriptorInit(h, &dC, m, n, n, 16, dt, CUSPARSE_ORDER_ROW));
    CKS(cusparseLtMatmulDescriptorInit(h, &mm, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                       CUSPARSE_OPERATION_NON_TRANSPOSE, &dA, &dB, &dC, &dC,
                                       CUSPARSE_COMPUTE_32F));
    CKS(cusparseLtMatmulAlgSelectionInit(h, &alg, &mm, CUSPARSELT_MATMUL_ALG_DEFAULT));
    CKS(cusparseLtMatmulPlanInit(h, &plan, &mm, &alg));
    CKS(cusparseLtMatmulGetWorkspace(h, &plan, &ws));
    return 0;
  }
  void destroy() {
    cusparseLtMatmulPlanDestroy(&plan);
    cusparseLtMatDescriptorDestroy(&dA);
    cusparseLtMatDescriptorDestroy(&dB);
    cusparseLtMatDescriptorDestroy(&dC);
  }
};

static int golden(int argc, char** argv) {
  if (argc != 7) return 2;
  int64_t m = std::atoll(argv[2]), k = std::atoll(argv[3]), n = std::atoll(argv[4]);
  const bool tile = std::string(argv[5]) == "tile";
  std::vector<__half> hA(m * k), hB(k * n);
  for (int64_t i = 0; i < m * k; ++i) hA[i] = __float2half(gen(1, i));
  for (int64_t i = 0; i < k * n; ++i) hB[i] = __float2half(gen(2, i));
  __half *A, *B, *C;
  int* d_valid;
  CK(cudaMalloc(&A, m * k * 2));
  CK(cudaMalloc(&B, k * n * 2));
  CK(cudaMalloc(&C, m * n * 2));
  CK(cudaMalloc(&d_valid, 4));
  CK(cudaMemcpy(A, hA.data(), m * k * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, hB.data(), k * n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(C, 0, m * n * 2));
  cusparseLtHandle_t h;
  CKS(cusparseLtInit(&h));
  Problem p;
  if (int rc = p.init(&h, m, n, k)) return rc;
  CKS(cusparseLtSpMMAPrune(&h, &p.mm, A, A, tile ? CUSPARSELT_PRUNE_SPMMA_TILE : CUSPARSELT_PRUNE_SPMMA_STRIP, 0));
  CKS(cusparseLtSpMMAPruneCheck(&h, &p.mm, A, d_valid, 0));
  int valid = -1;
  CK(cudaMemcpy(&valid, d_valid, 4, cudaMemcpyDeviceToHost));
  size_t csz = 0, cbuf = 0;
  CKS(cusparseLtSpMMACompressedSize(&h, &p.plan, &csz, &cbuf));
  void *Ac, *Abuf = nullptr, *ws = nullptr;
  CK(cudaMalloc(&Ac, csz));
  if (cbuf) CK(cudaMalloc(&Abuf, cbuf));
  if (p.ws) CK(cudaMalloc(&ws, p.ws));
  CKS(cusparseLtSpMMACompress(&h, &p.plan, A, Ac, Abuf, 0));
  float alpha = 1.f, beta = 0.f;
  CKS(cusparseLtMatmul(&h, &p.plan, &alpha, Ac, B, &beta, C, C, ws, nullptr, 0));
  CK(cudaDeviceSynchronize());
  std::vector<__half> hP(m * k), hD(m * n);
  CK(cudaMemcpy(hP.data(), A, m * k * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hD.data(), C, m * n * 2, cudaMemcpyDeviceToHost));
  FILE* f = std::fopen(argv[6], "wb");
  if (!f) return 5;
  int64_t hdr[5] = {m, k, n, tile ? 0 : 1, valid};
  std::fwrite(hdr, 8, 5, f);
  std::fwrite(hA.data(), 2, m * k, f);
  std::fwrite(hP.data(), 2, m * k, f);
  std::fwrite(hB.data(), 2, k * n, f);
  std::fwrite(hD.data(), 2, m * n, f);
  std::fclose(f);
  std::printf("golden %lldx%lldx%lld %s: PruneCheck=%d compressed=%zu bytes\n", (long long)m, (long long)k,
              (long long)n, argv[5], valid, csz);
  p.destroy();
  cusparseLtDestroy(&h);
  return 0;
}

// prunefile <m> <k> <tile|strip> <f16|bf16> <in.bin> <out.bin>: cusparseLtSpMMAPrune (spmma.hxx:86) on the
// 16-bit matrix stored in in.bin (m*k uint16, row-major); writes the pruned matrix.  Used by
// tests/golden/make_tile_probe.py to learn how the closed library breaks ties between 4x4 patterns.
static int prunefile(int argc, char** argv) {
  if (argc != 8) return 2;
  int64_t m = std::atoll(argv[2]), k = std::atoll(argv[3]);
  const bool tile = std::string(argv[4]) == "tile";
  const cudaDataType dt = std::string(argv[5]) == "bf16" ? CUDA_R_16BF : CUDA_R_16F;
  std::vector<uint16_t> hA(m * k);
  FILE* f = std::fopen(argv[6], "rb");
  if (!f || std::fread(hA.data(), 2, m * k, f) != (size_t)(m * k)) return 5;
  std::fclose(f);
  void* A;
  CK(cudaMalloc(&A, m * k * 2));
  CK(cudaMemcpy(A, hA.data(), m * k * 2, cudaMemcpyHostToDevice));
  cusparseLtHandle_t h;
  CKS(cusparseLtInit(&h));
  Problem p;
  if (int rc = p.init(&h, m, 64, k, dt)) return rc;
  CKS(cusparseLtSpMMAPrune(&h, &p.mm, A, A, tile ? CUSPARSELT_PRUNE_SPMMA_TILE : CUSPARSELT_PRUNE_SPMMA_STRIP, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hA.data(), A, m * k * 2, cudaMemcpyDeviceToHost));
  f = std::fopen(argv[7], "wb");
  if (!f) return 5;
  std::fwrite(hA.data(), 2, m * k, f);
  std::fclose(f);
  p.destroy();
  cusparseLtDestroy(&h);
  return 0;
}

static float median(std::vector<float> v) {
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

static int sweep(int argc, char** argv) {
  if (argc != 4) return 2;
  const int batch = std::atoi(argv[3]);
  std::ifstream in(argv[2]);
  if (!in.is_open()) {
    std::fprintf(stderr, "cannot open %s\n", argv[2]);
    return 2;
  }
  std::string line;
  std::getline(in, line);
  struct Shape { int64_t M, N, K; };
  std::vector<Shape> shapes;
  while (std::getline(in, line)) {
    long long v[4];
    if (std::sscanf(line.c_str(), "%lld,%lld,%lld,%lld", &v[0], &v[1], &v[2], &v[3]) == 4)
      shapes.push_back({v[1], v[0] * batch, v[2]});
  }
  cusparseLtHandle_t h;
  CKS(cusparseLtInit(&h));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  void* flush;
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  int* d_valid;
  CK(cudaMalloc(&d_valid, 4));
  double tot_prune = 0, tot_comp = 0, tot_mm = 0, flops = 0, bytes = 0;
  int skipped = 0;
  for (const Shape& s : shapes) {
    // cusparseLt wants k % 16 == 0 for fp16 structured operands: pad K like we do (zero columns)
    const int64_t K = (s.K + 15) / 16 * 16, M = s.M, N = s.N;
    __half *A0, *A, *B, *C;
    CK(cudaMalloc(&A0, M * K * 2));
    CK(cudaMalloc(&A, M * K * 2));
    CK(cudaMalloc(&B, K * N * 2));
    CK(cudaMalloc(&C, M * N * 2));
    {
      std::vector<__half> hA(M * K);
      for (int64_t i = 0; i < M * K; ++i) hA[i] = __float2half((i % K) < s.K ? gen(1, i) : 0.f);
      CK(cudaMemcpy(A0, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
      CK(cudaMemset(B, 0x3c, K * N * 2));  // fp16 1.0586: timing does not depend on the values
    }
    Problem p;
    if (p.init(&h, M, N, K)) {
      ++skipped;
      continue;
    }
    size_t csz = 0, cbuf = 0;
    CKS(cusparseLtSpMMACompressedSize(&h, &p.plan, &csz, &cbuf));
    void *Ac, *Abuf = nullptr, *ws = nullptr;
    CK(cudaMalloc(&Ac, csz));
    if (cbuf) CK(cudaMalloc(&Abuf, cbuf));
    if (p.ws) CK(cudaMalloc(&ws, p.ws));
    float alpha = 1.f, beta = 0.f;
    std::vector<float> tp, tc, tm;
    for (int it = 0; it < 7; ++it) {
      float ms;
      CK(cudaMemcpy(A, A0, M * K * 2, cudaMemcpyDeviceToDevice));
      CK(cudaMemsetAsync(flush, it, flush_bytes, 0));
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtSpMMAPrune(&h, &p.mm, A, A, CUSPARSELT_PRUNE_SPMMA_STRIP, 0));
      CKS(cusparseLtSpMMAPruneCheck(&h, &p.mm, A, d_valid, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tp.push_back(ms);
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtSpMMACompress(&h, &p.plan, A, Ac, Abuf, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tc.push_back(ms);
      CK(cudaMemsetAsync(flush, it + 1, flush_bytes, 0));
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtMatmul(&h, &p.plan, &alpha, Ac, B, &beta, C, C, ws, nullptr, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tm.push_back(ms);
    }
    const float pm = median(tp), cm = median(tc), mm = median(tm);
    std::printf("layer M=%lld K=%lld N=%lld prune %.4f ms compress %.4f ms matmul %.4f ms (%.1f TFLOP/s)\n",
                (long long)M, (long long)s.K, (long long)N, pm, cm, mm, 2.0 * M * N * s.K / mm / 1e9);
    tot_prune += pm;
    tot_comp += cm;
    tot_mm += mm;
    flops += 2.0 * M * N * s.K;
    bytes += 2.0 * s.K * N + 2.0 * M * N + 1.125 * M * s.K;
    p.destroy();
    cudaFree(A0); cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Ac);
    if (Abuf) cudaFree(Abuf);
    if (ws) cudaFree(ws);
  }
  int ver = 0;
  cusparseLtGetProperty(MAJOR_VERSION, &ver);
  int minor = 0, patch = 0;
  cusparseLtGetProperty(MINOR_VERSION, &minor);
  cusparseLtGetProperty(PATCH_LEVEL, &patch);
  std::printf("{\"library\": \"cusparseLt %d.%d.%d\", \"layers\": %zu, \"skipped\": %d, \"dtype\": \"f16\", "
              "\"compute\": \"32F\", \"timing\": \"median of 5 single launches, L2 flushed before each\", "
              "\"prune_ms\": %.4f, \"compress_ms\": %.4f, \"matmul_ms\": %.4f, "
              "\"matmul_tflops\": %.2f, \"matmul_gbs\": %.1f, \"total_tflops\": %.2f}\n",
              ver, minor, patch, shapes.size(), skipped, tot_prune, tot_comp, tot_mm, flops / tot_mm / 1e9,
              bytes / tot_mm / 1e6, flops / (tot_mm + tot_prune + tot_comp) / 1e9);
  cusparseLtDestroy(&h);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "golden") return golden(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "sweep") return sweep(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "prunefile") return prunefile(argc, argv);
  std::fprintf(stderr, "usage: %s golden m k n tile|strip out.bin | sweep shapes.csv batch\n", argv[0]);
  return 2;
}

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

// device-side random fill (values in [-1, 1) like `gen`, any count): timing must not run on a constant operand
__global__ void fill_random_half(__half* p, size_t n, uint64_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t z = (seed * 0x632BE59BD9B4E019ull + i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = __float2half((float)((int64_t)(z >> 40) - (1 << 23)) / (float)(1 << 23));
  }
}

// cusparselt_ref sweep <table.csv> <batch> : the reference's spmma call sequence (include/sparsify.me/spmma.hxx:51-114)
// over a whole shape table in the weights orientation, as a like-for-like comparator for bench.py:
//   * every layer keeps its own A / B / C (nothing is re-read from L2 between layers), B is random;
//   * the kernel of every plan is chosen by cusparseLtMatmulSearch (the library's own auto-tuner), not ALG_DEFAULT;
//   * two timings of the matmuls: each layer alone (median of 5 launches, L2 flushed before each, one event pair per
//     launch) and the whole table BACK TO BACK (all layers enqueued one after the other, 20 steps under one event
//     pair) -- the mode bench.py times our plan in.
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
  struct Layer {
    Shape s;
    int64_t K;
    Problem p;
    __half *A0 = nullptr, *A = nullptr, *B = nullptr, *C = nullptr;
    void *Ac = nullptr, *Abuf = nullptr, *ws = nullptr;
    float prune_ms = 0, comp_ms = 0, mm_ms = 0;
  };
  std::vector<Layer*> layers;
  const float alpha = 1.f, beta = 0.f;
  int skipped = 0, searched = 0;
  for (size_t li = 0; li < shapes.size(); ++li) {
    const Shape& s = shapes[li];
    Layer* L = new Layer();
    L->s = s;
    // cusparseLt wants k % 16 == 0 for fp16 structured operands: pad K like we do (zero columns)
    const int64_t K = L->K = (s.K + 15) / 16 * 16, M = s.M, N = s.N;
    CK(cudaMalloc(&L->A0, M * K * 2));
    CK(cudaMalloc(&L->A, M * K * 2));
    CK(cudaMalloc(&L->B, K * N * 2));
    CK(cudaMalloc(&L->C, M * N * 2));
    {
      std::vector<__half> hA(M * K);
      for (int64_t i = 0; i < M * K; ++i) hA[i] = __float2half((i % K) < s.K ? gen(1 + li, i) : 0.f);
      CK(cudaMemcpy(L->A0, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
      fill_random_half<<<1184, 256>>>(L->B, (size_t)K * N, 1000 + li);
      CK(cudaGetLastError());
    }
    if (L->p.init(&h, M, N, K)) {
      ++skipped;
      delete L;
      continue;
    }
    size_t csz = 0, cbuf = 0;
    CKS(cusparseLtSpMMACompressedSize(&h, &L->p.plan, &csz, &cbuf));
    CK(cudaMalloc(&L->Ac, csz));
    if (cbuf) CK(cudaMalloc(&L->Abuf, cbuf));
    // prune + compress, timed like the reference's phases (spmma.hxx:82-104)
    std::vector<float> tp, tc;
    for (int it = 0; it < 5; ++it) {
      float ms;
      CK(cudaMemcpy(L->A, L->A0, M * K * 2, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtSpMMAPrune(&h, &L->p.mm, L->A, L->A, CUSPARSELT_PRUNE_SPMMA_STRIP, 0));
      CKS(cusparseLtSpMMAPruneCheck(&h, &L->p.mm, L->A, d_valid, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tp.push_back(ms);
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtSpMMACompress(&h, &L->p.plan, L->A, L->Ac, L->Abuf, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tc.push_back(ms);
    }
    L->prune_ms = median(tp);
    L->comp_ms = median(tc);
    // the library's auto-tuner picks the kernel (it may need a larger workspace than the default algorithm)
    {
      size_t ws_max = std::max<size_t>(L->p.ws, 64ull << 20);
      CK(cudaMalloc(&L->ws, ws_max));
      cusparseStatus_t st = cusparseLtMatmulSearch(&h, &L->p.plan, &alpha, L->Ac, L->B, &beta, L->C, L->C, L->ws, nullptr, 0);
      if (st == CUSPARSE_STATUS_SUCCESS) {
        ++searched;
        size_t need = 0;
        CKS(cusparseLtMatmulGetWorkspace(&h, &L->p.plan, &need));
        if (need > ws_max) {
          cudaFree(L->ws);
          CK(cudaMalloc(&L->ws, need));
        }
      }
    }
    layers.push_back(L);
  }
  // ---- each layer alone, L2 flushed before every launch ----
  double tot_prune = 0, tot_comp = 0, tot_mm = 0, flops = 0, bytes = 0;
  for (Layer* L : layers) {
    std::vector<float> tm;
    for (int it = 0; it < 7; ++it) {
      float ms;
      CK(cudaMemsetAsync(flush, it + 1, flush_bytes, 0));
      CK(cudaEventRecord(e0, 0));
      CKS(cusparseLtMatmul(&h, &L->p.plan, &alpha, L->Ac, L->B, &beta, L->C, L->C, L->ws, nullptr, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) tm.push_back(ms);
    }
    L->mm_ms = median(tm);
    const Shape& s = L->s;
    std::printf("layer M=%lld K=%lld N=%lld prune %.4f ms compress %.4f ms matmul %.4f ms (%.1f TFLOP/s)\n", (long long)s.M,
                (long long)s.K, (long long)s.N, L->prune_ms, L->comp_ms, L->mm_ms, 2.0 * s.M * s.N * s.K / L->mm_ms / 1e9);
    tot_prune += L->prune_ms;
    tot_comp += L->comp_ms;
    tot_mm += L->mm_ms;
    flops += 2.0 * s.M * s.N * s.K;
    bytes += 2.0 * s.K * s.N + 2.0 * s.M * s.N + 1.125 * s.M * s.K;
  }
  // ---- the whole table back to back: 3 warm-up steps, 20 timed steps under one event pair ----
  float b2b_ms = 0;
  {
    const int steps = 20;
    for (int it = 0; it < 3 + steps; ++it) {
      if (it == 3) CK(cudaEventRecord(e0, 0));
      for (Layer* L : layers)
        CKS(cusparseLtMatmul(&h, &L->p.plan, &alpha, L->Ac, L->B, &beta, L->C, L->C, L->ws, nullptr, 0));
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&b2b_ms, e0, e1));
    b2b_ms /= steps;
  }
  int ver = 0;
  cusparseLtGetProperty(MAJOR_VERSION, &ver);
  int minor = 0, patch = 0;
  cusparseLtGetProperty(MINOR_VERSION, &minor);
  cusparseLtGetProperty(PATCH_LEVEL, &patch);
  std::printf("{\"library\": \"cusparseLt %d.%d.%d\", \"layers\": %zu, \"skipped\": %d, \"dtype\": \"f16\", "
              "\"compute\": \"32F\", \"search\": %s, \"searched_layers\": %d, \"operands\": \"per-layer buffers, random B\", "
              "\"timing\": \"single_ms: sum over layers of the median of 5 launches, L2 flushed before each; "
              "back_to_back_ms: all layers enqueued back to back, mean of 20 steps under one event pair\", "
              "\"prune_ms\": %.4f, \"compress_ms\": %.4f, \"matmul_ms\": %.4f, \"single_ms\": %.4f, \"back_to_back_ms\": %.4f, "
              "\"matmul_tflops\": %.2f, \"matmul_gbs\": %.1f, \"back_to_back_tflops\": %.2f, \"total_tflops\": %.2f}\n",
              ver, minor, patch, shapes.size(), skipped, searched == (int)layers.size() ? "true" : "false", searched,
              tot_prune, tot_comp, tot_mm, tot_mm, b2b_ms, flops / tot_mm / 1e9, bytes / tot_mm / 1e6, flops / b2b_ms / 1e9,
              flops / (tot_mm + tot_prune + tot_comp) / 1e9);
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

// cusparse_ref.cu -- golden generator for the unstructured path: issues THE cuSPARSE generic-API call
// sequence the reference intends at include/sparsify.me/spmm.hxx:164-187 (batched::strided_coo) and
// :57-67,107-110 (batched::spmm, blocked-ELL) against the on-box cuSPARSE 12.x.
// TEST INFRASTRUCTURE ONLY; never linked into the product.
//
// The reference's own drivers do not compile at HEAD (spmm.hxx:172,175 undeclared identifiers) and carry
// the runtime defects listed in SURVEY.md 8a (host pointer as workspace, `type_t**` as values pointer, ...),
// so this is OUR driver around the library the reference delegates the arithmetic to:
//   COO : CreateCoo(m, k, nnz, rows, cols, vals, 32I, base 0, R_32F) -> CooSetStridedBatch(nb, 0)
//         -> CreateDnMat(B k x n ld k COL) + DnMatSetStridedBatch(nb, ld*n) -> same for C (m x n ld m)
//         -> SpMM_bufferSize -> SpMM(N, N, alpha, A, B, beta, C, R_32F, CUSPARSE_SPMM_COO_ALG4)
//   BELL: per batch element CreateBlockedEll(m, k, block, ell_cols, colInd 32I, values R_32F)
//         -> CreateDnMat(B k x n ld k COL), CreateDnMat(C_b m x n ld m COL) -> SpMM(ALG_DEFAULT)
//
//   cusparse_ref coo  <m> <k> <n> <nb> <threshold> <alpha> <beta> <out.bin>
//   cusparse_ref bell <m> <k> <n> <nb> <block> <out.bin>
// Inputs come from the splitmix64 counter generator `gen` (multiples of 1/64 in [-1, 1): every product
// and partial sum is exact in fp32, so the library result, our kernels and the fp64 oracle must agree
// BIT FOR BIT whatever the summation order).  Output file: int64 header, then the arrays named below.
#include <cuda_runtime.h>
#include <cusparse.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <array>
#include <map>
#include <vector>

#define CK(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
      std::exit(3);                                                                                 \
    }                                                                                               \
  } while (0)
#define CKS(x)                                                                                      \
  do {                                                                                              \
    cusparseStatus_t s_ = (x);                                                                      \
    if (s_ != CUSPARSE_STATUS_SUCCESS) {                                                            \
      std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cusparseGetErrorString(s_)); \
      return 4;                                                                                     \
    }                                                                                               \
  } while (0)

// element i of tensor t; same formula in tests/golden/make_golden.py and oracle/cusparselt_ref.cu
static inline float gen(uint64_t t, uint64_t i) {
  uint64_t z = (t * 0x632BE59BD9B4E019ull + i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  int q = (int)(z >> 57) - 64;  // [-64, 63]
  return (float)q / 64.0f;
}

template <typename T>
static T* to_dev(const std::vector<T>& h) {
  T* d = nullptr;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static void dump(const char* path, const std::vector<int64_t>& hdr, const std::vector<std::pair<const void*, size_t>>& parts) {
  FILE* f = std::fopen(path, "wb");
  if (!f) { std::perror(path); std::exit(5); }
  std::fwrite(hdr.data(), 8, hdr.size(), f);
  for (auto& p : parts) std::fwrite(p.first, 1, p.second, f);
  std::fclose(f);
}

static int run_coo(int argc, char** argv) {
  if (argc < 10) return 2;
  const int64_t m = atoll(argv[2]), k = atoll(argv[3]), n = atoll(argv[4]), nb = atoll(argv[5]);
  const float thr = (float)atof(argv[6]), alpha = (float)atof(argv[7]), beta = (float)atof(argv[8]);
  std::vector<int32_t> rows, cols;
  std::vector<float> vals;
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < k; ++j) {
      const float a = gen(1, (uint64_t)(i * k + j));
      if (std::fabs(a) > thr) { rows.push_back((int32_t)i); cols.push_back((int32_t)j); vals.push_back(a); }
    }
  const int64_t nnz = (int64_t)vals.size();
  std::vector<float> B((size_t)nb * n * k), C((size_t)nb * n * m);
  for (size_t i = 0; i < B.size(); ++i) B[i] = gen(2, i);
  for (size_t i = 0; i < C.size(); ++i) C[i] = gen(3, i);
  int32_t *dr = to_dev(rows), *dc = to_dev(cols);
  float *dv = to_dev(vals), *dB = to_dev(B), *dC = to_dev(C);
  cusparseHandle_t h;
  CKS(cusparseCreate(&h));
  cusparseSpMatDescr_t A;
  cusparseDnMatDescr_t mB, mC;
  CKS(cusparseCreateCoo(&A, m, k, nnz, dr, dc, dv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
  CKS(cusparseCooSetStridedBatch(A, (int)nb, 0));
  CKS(cusparseCreateDnMat(&mB, k, n, k, dB, CUDA_R_32F, CUSPARSE_ORDER_COL));
  CKS(cusparseDnMatSetStridedBatch(mB, (int)nb, k * n));
  CKS(cusparseCreateDnMat(&mC, m, n, m, dC, CUDA_R_32F, CUSPARSE_ORDER_COL));
  CKS(cusparseDnMatSetStridedBatch(mC, (int)nb, m * n));
  size_t bufsz = 0;
  CKS(cusparseSpMM_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB,
                              &beta, mC, CUDA_R_32F, CUSPARSE_SPMM_COO_ALG4, &bufsz));
  void* buf = nullptr;
  CK(cudaMalloc(&buf, bufsz + 16));
  CKS(cusparseSpMM(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB, &beta, mC,
                   CUDA_R_32F, CUSPARSE_SPMM_COO_ALG4, buf));
  CK(cudaDeviceSynchronize());
  std::vector<float> out(C.size());
  CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
  // header{m,k,n,nb,nnz}; rows[nnz] i32, cols[nnz] i32, vals[nnz] f32, C_out[nb*n*m] f32
  dump(argv[9], {m, k, n, nb, nnz},
       {{rows.data(), rows.size() * 4}, {cols.data(), cols.size() * 4}, {vals.data(), vals.size() * 4},
        {out.data(), out.size() * 4}});
  std::printf("coo m=%lld k=%lld n=%lld nb=%lld nnz=%lld ok\n", (long long)m, (long long)k, (long long)n,
              (long long)nb, (long long)nnz);
  return 0;
}

static int run_bell(int argc, char** argv) {
  if (argc < 8) return 2;
  const int64_t m = atoll(argv[2]), k = atoll(argv[3]), n = atoll(argv[4]), nb = atoll(argv[5]), block = atoll(argv[6]);
  const int64_t ell_cols = k / 2, bcols = ell_cols / block, brows = m / block;
  // the reference driver's construction (examples/spmm.cu:45-84): every block row keeps bcols distinct
  // block columns, ascending; here: block column j of block row i = (i + 2*j) mod (k/block), sorted
  std::vector<int32_t> ci32((size_t)nb * brows * bcols);
  std::vector<int64_t> ci64(ci32.size());
  for (int64_t b = 0; b < nb; ++b)
    for (int64_t i = 0; i < brows; ++i) {
      std::vector<int64_t> row(bcols);
      for (int64_t j = 0; j < bcols; ++j) row[j] = (i + b + 2 * j) % (k / block);
      std::sort(row.begin(), row.end());
      for (int64_t j = 0; j < bcols; ++j) {
        ci32[(b * brows + i) * bcols + j] = (int32_t)row[j];
        ci64[(b * brows + i) * bcols + j] = row[j];
      }
    }
  std::vector<float> V((size_t)nb * m * ell_cols), B((size_t)n * k), out((size_t)nb * n * m);
  for (size_t i = 0; i < V.size(); ++i) V[i] = gen(4, i);
  for (size_t i = 0; i < B.size(); ++i) B[i] = gen(5, i);
  int32_t* dci = to_dev(ci32);
  float *dV = to_dev(V), *dB = to_dev(B), *dC = nullptr;
  CK(cudaMalloc(&dC, out.size() * 4));
  CK(cudaMemset(dC, 0, out.size() * 4));
  cusparseHandle_t h;
  CKS(cusparseCreate(&h));
  const float alpha = 1.f, beta = 0.f;
  for (int64_t b = 0; b < nb; ++b) {
    cusparseSpMatDescr_t A;
    cusparseDnMatDescr_t mB, mC;
    CKS(cusparseCreateBlockedEll(&A, m, k, block, ell_cols, dci + b * brows * bcols, dV + b * m * ell_cols,
                                 CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    CKS(cusparseCreateDnMat(&mB, k, n, k, dB, CUDA_R_32F, CUSPARSE_ORDER_COL));
    CKS(cusparseCreateDnMat(&mC, m, n, m, dC + b * n * m, CUDA_R_32F, CUSPARSE_ORDER_COL));
    size_t bufsz = 0;
    CKS(cusparseSpMM_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB,
                                &beta, mC, CUDA_R_32F, CUSPARSE_SPMM_ALG_DEFAULT, &bufsz));
    void* buf = nullptr;
    CK(cudaMalloc(&buf, bufsz + 16));
    CKS(cusparseSpMM(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB, &beta, mC,
                     CUDA_R_32F, CUSPARSE_SPMM_ALG_DEFAULT, buf));
    CK(cudaDeviceSynchronize());
    CK(cudaFree(buf));
  }
  CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
  // header{m,k,n,nb,block,ell_cols}; col_idx[nb*brows*bcols] i64, C_out[nb*n*m] f32
  dump(argv[7], {m, k, n, nb, block, ell_cols}, {{ci64.data(), ci64.size() * 8}, {out.data(), out.size() * 4}});
  std::printf("bell m=%lld k=%lld n=%lld nb=%lld block=%lld ok\n", (long long)m, (long long)k, (long long)n,
              (long long)nb, (long long)block);
  return 0;
}

// cusparse_ref time <m> <k> <n> <nb> <keep_fraction> : median-of-5 time of the batched COO SpMM (ALG4) on random
// fp32 data with ~keep_fraction of A kept -- the comparator column of tools/spmm_sweep.py
static int run_time(int argc, char** argv) {
  if (argc < 7) return 2;
  const int64_t m = atoll(argv[2]), k = atoll(argv[3]), n = atoll(argv[4]), nb = atoll(argv[5]);
  const double keep = atof(argv[6]);
  std::vector<int32_t> rows, cols;
  std::vector<float> vals;
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < k; ++j)
      if (rnd() < keep) { rows.push_back((int32_t)i); cols.push_back((int32_t)j); vals.push_back((float)(2 * rnd() - 1)); }
  const int64_t nnz = (int64_t)vals.size();
  int32_t *dr = to_dev(rows), *dc = to_dev(cols);
  float* dv = to_dev(vals);
  float *dB = nullptr, *dC = nullptr;
  CK(cudaMalloc(&dB, (size_t)nb * n * k * 4));
  CK(cudaMalloc(&dC, (size_t)nb * n * m * 4));
  CK(cudaMemset(dB, 0x3c, (size_t)nb * n * k * 4));
  CK(cudaMemset(dC, 0, (size_t)nb * n * m * 4));
  cusparseHandle_t h;
  CKS(cusparseCreate(&h));
  cusparseSpMatDescr_t A;
  cusparseDnMatDescr_t mB, mC;
  const float alpha = 1.f, beta = 0.f;
  CKS(cusparseCreateCoo(&A, m, k, nnz, dr, dc, dv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
  CKS(cusparseCooSetStridedBatch(A, (int)nb, 0));
  CKS(cusparseCreateDnMat(&mB, k, n, k, dB, CUDA_R_32F, CUSPARSE_ORDER_COL));
  CKS(cusparseDnMatSetStridedBatch(mB, (int)nb, k * n));
  CKS(cusparseCreateDnMat(&mC, m, n, m, dC, CUDA_R_32F, CUSPARSE_ORDER_COL));
  CKS(cusparseDnMatSetStridedBatch(mC, (int)nb, m * n));
  size_t bufsz = 0;
  CKS(cusparseSpMM_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB,
                              &beta, mC, CUDA_R_32F, CUSPARSE_SPMM_COO_ALG4, &bufsz));
  void* buf = nullptr;
  CK(cudaMalloc(&buf, bufsz + 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ts;
  for (int it = 0; it < 7; ++it) {
    CK(cudaEventRecord(e0));
    CKS(cusparseSpMM(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, mB, &beta, mC,
                     CUDA_R_32F, CUSPARSE_SPMM_COO_ALG4, buf));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= 2) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  std::printf("{\"library\": \"cuSPARSE COO_ALG4\", \"m\": %lld, \"k\": %lld, \"n\": %lld, \"nb\": %lld, \"nnz\": %lld, \"us\": %.1f}\n",
              (long long)m, (long long)k, (long long)n, (long long)nb, (long long)nnz, ts[ts.size() / 2] * 1e3);
  return 0;
}

// cusparse_ref timebell <csv> [batch] : the reference's batched::spmm as its driver runs it over a shape table
// (examples/spmm.cu:40-116, include/sparsify.me/spmm.hxx:57-115): per CSV row (m, n, k, b) b blocked-ELL matrices
// (block 2, ell_cols = k/2, ascending unique block-column ids), ONE shared B, one cusparseSpMM per batch element,
// each on its own stream; timed like the reference's timer (events on the null stream around the fan-out), best
// of 3 after a warm-up.  One CSV line per row and a JSON summary: the same-box comparator for our `spmm` column.
static int run_timebell(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[2], "r");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
  const int64_t batch_override = argc > 3 ? atoll(argv[3]) : 0;
  char line[256];
  std::vector<std::array<int64_t, 4>> shapes;
  bool first = true;
  while (std::fgets(line, sizeof(line), f)) {
    if (first) { first = false; continue; }
    long long m, n, k, b;
    if (std::sscanf(line, "%lld,%lld,%lld,%lld", &m, &n, &k, &b) == 4) shapes.push_back({m, n, k, batch_override ? batch_override : b});
  }
  std::fclose(f);
  cusparseHandle_t h;
  CKS(cusparseCreate(&h));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::map<std::array<int64_t, 4>, double> memo;
  double total_ms = 0;
  std::printf("m,n,k,b,cusparse_bell_ms\n");
  for (const auto& sh : shapes) {
    if (memo.count(sh)) { total_ms += memo[sh]; std::printf("%lld,%lld,%lld,%lld,%.5f\n", (long long)sh[0], (long long)sh[1], (long long)sh[2], (long long)sh[3], memo[sh]); continue; }
    const int64_t m = sh[0], n = sh[1], k = sh[2], nb = sh[3], block = 2;
    const int64_t ell_cols = k / 2 / block * block, bcols = ell_cols / block, brows = (m + block - 1) / block, nbc = k / block;
    // cuSPARSE rejects rows / cols that are not multiples of the block (k = 147)
    if (bcols == 0 || m % block || k % block) { std::printf("%lld,%lld,%lld,%lld,nan\n", (long long)m, (long long)n, (long long)k, (long long)nb); continue; }
    std::vector<int32_t> ci((size_t)nb * brows * bcols);
    for (int64_t b = 0; b < nb; ++b)
      for (int64_t i = 0; i < brows; ++i) {
        const int64_t phase = (i + b) & 1;  // every other block column: ascending, unique
        for (int64_t j = 0; j < bcols; ++j) ci[(b * brows + i) * bcols + j] = (int32_t)std::min<int64_t>(2 * j + phase, nbc - 1);
      }
    int32_t* dci = to_dev(ci);
    float *dV = nullptr, *dB = nullptr, *dC = nullptr;
    CK(cudaMalloc(&dV, (size_t)nb * m * ell_cols * 4));
    CK(cudaMalloc(&dB, (size_t)n * k * 4));
    CK(cudaMalloc(&dC, (size_t)nb * n * m * 4));
    CK(cudaMemset(dV, 0x3c, (size_t)nb * m * ell_cols * 4));
    CK(cudaMemset(dB, 0x3c, (size_t)n * k * 4));
    const float alpha = 1.f, beta = 0.f;
    std::vector<cusparseSpMatDescr_t> A(nb);
    std::vector<cusparseDnMatDescr_t> mC(nb);
    std::vector<cudaStream_t> st(nb);
    std::vector<void*> bufs(nb, nullptr);
    cusparseDnMatDescr_t mB;
    CKS(cusparseCreateDnMat(&mB, k, n, k, dB, CUDA_R_32F, CUSPARSE_ORDER_COL));
    for (int64_t b = 0; b < nb; ++b) {
      CKS(cusparseCreateBlockedEll(&A[b], m, k, block, ell_cols, dci + b * brows * bcols, dV + b * m * ell_cols,
                                   CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
      CKS(cusparseCreateDnMat(&mC[b], m, n, m, dC + b * n * m, CUDA_R_32F, CUSPARSE_ORDER_COL));
      size_t bufsz = 0;
      CKS(cusparseSpMM_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A[b], mB,
                                  &beta, mC[b], CUDA_R_32F, CUSPARSE_SPMM_ALG_DEFAULT, &bufsz));
      CK(cudaMalloc(&bufs[b], bufsz + 16));
      CK(cudaStreamCreate(&st[b]));
    }
    double best = 1e30;
    for (int it = 0; it < 4; ++it) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, 0));
      for (int64_t b = 0; b < nb; ++b) {
        CKS(cusparseSetStream(h, st[b]));
        CKS(cusparseSpMM(h, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A[b], mB, &beta,
                         mC[b], CUDA_R_32F, CUSPARSE_SPMM_ALG_DEFAULT, bufs[b]));
      }
      CK(cudaEventRecord(e1, 0));  // the null stream waits for every blocking stream, like the reference's timer
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 1 && ms < best) best = ms;
    }
    CKS(cusparseSetStream(h, 0));
    for (int64_t b = 0; b < nb; ++b) {
      cusparseDestroySpMat(A[b]);
      cusparseDestroyDnMat(mC[b]);
      cudaFree(bufs[b]);
      cudaStreamDestroy(st[b]);
    }
    cusparseDestroyDnMat(mB);
    cudaFree(dci); cudaFree(dV); cudaFree(dB); cudaFree(dC);
    memo[sh] = best;
    total_ms += best;
    std::printf("%lld,%lld,%lld,%lld,%.5f\n", (long long)m, (long long)n, (long long)k, (long long)nb, best);
    std::fflush(stdout);
  }
  std::printf("{\"library\": \"cuSPARSE blocked-ELL SpMM (block 2, ell_cols k/2, fp32, one call per batch element on its own stream)\", "
              "\"rows\": %zu, \"sum_ms\": %.4f}\n", shapes.size(), total_ms);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "timebell") return run_timebell(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "time") return run_time(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "coo") return run_coo(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "bell") return run_bell(argc, argv);
  std::fprintf(stderr, "usage: cusparse_ref coo <m> <k> <n> <nb> <thr> <alpha> <beta> <out.bin>\n"
                       "       cusparse_ref bell <m> <k> <n> <nb> <block> <out.bin>\n"
                       "       cusparse_ref time <m> <k> <n> <nb> <keep>\n"
                       "       cusparse_ref timebell <table.csv> [batch]\n");
  return 2;
}

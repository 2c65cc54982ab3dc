// header_parity.cu -- test driver: calls the five operator TEMPLATES of include/sparsify.me exactly as the
// reference's drivers do (examples/sparsify.cu:46, spmma.cu:61-62, spmm.cu:115-116, batched_coo.cu:102-111,
// gemm.cu:93-95), on deterministic inputs, and dumps every input and output as raw little-endian arrays into a
// directory.  tests/test_gpu_headers.py runs it on the GPU box and checks the dumps against the CPU oracle, so
// the C++ header layer -- not only the C ABI underneath it -- is parity-tested.
//
//     header_parity <out_dir>
//
// Built by __graft_entry__.build() into tests/cpp/bin/ (nvcc cross-compiles without a GPU).
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/containers/ell.hxx>
#include <sparsify.me/gemm.hxx>
#include <sparsify.me/sparsify.hxx>
#include <sparsify.me/spmm.hxx>
#include <sparsify.me/spmma.hxx>
#include <sparsify.me/util/util.hxx>

namespace {

std::string g_dir;

template <typename T>
void dump(const std::string& name, const T* p, std::size_t n) {
  std::ofstream f(g_dir + "/" + name + ".bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), (std::streamsize)(n * sizeof(T)));
}
template <typename T>
void dump(const std::string& name, const thrust::host_vector<T>& v) { dump(name, v.data(), v.size()); }
template <typename T>
void dump(const std::string& name, const thrust::device_vector<T>& v) {
  thrust::host_vector<T> h = v;
  dump(name, h.data(), h.size());
}

// splitmix64 counter generator
std::uint64_t mix(std::uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// U(-1, 1) with 24 random bits
float uniform(std::uint64_t stream, std::uint64_t i) {
  return (float)((mix(stream * 0x632BE59BD9B4E019ull + i) >> 40) + 0.5) / 8388608.0f - 1.0f;
}
thrust::host_vector<float> random_f32(std::uint64_t stream, std::size_t n) {
  thrust::host_vector<float> v(n);
  for (std::size_t i = 0; i < n; ++i) v[i] = uniform(stream, i);
  return v;
}
template <typename T>
thrust::host_vector<T> cast_to(const thrust::host_vector<float>& v) {
  thrust::host_vector<T> o(v.size());
  for (std::size_t i = 0; i < v.size(); ++i) o[i] = T(v[i]);
  return o;
}

void write_meta(const std::string& text) {
  std::ofstream f(g_dir + "/meta.txt", std::ios::app);
  f << text << "\n";
}

// ---- sparsify<2,2> (examples/sparsify.cu:46) ----
void case_sparsify(std::size_t m, std::size_t n) {
  thrust::host_vector<float> h = random_f32(1, m * n);
  thrust::device_vector<float> w = h;
  thrust::device_vector<std::size_t> mask(m * n);
  dump("sparsify_in", h);
  sparsifyme::sparsify<2, 2>(w.data().get(), mask.data().get(), m, n);
  cudaDeviceSynchronize();
  dump("sparsify_out", w);
  dump("sparsify_mask", mask);
  write_meta("sparsify " + std::to_string(m) + " " + std::to_string(n));
}

// ---- spmma<type_t> (examples/spmma.cu:61-62) ----
template <typename T>
void case_spmma(const char* tag, std::size_t m, std::size_t n, std::size_t k) {
  thrust::host_vector<T> hA = cast_to<T>(random_f32(2, m * k)), hB = cast_to<T>(random_f32(3, k * n));
  thrust::device_vector<T> A = hA, B = hB, C(m * n);
  dump(std::string("spmma_") + tag + "_a_in", hA);
  dump(std::string("spmma_") + tag + "_b", hB);
  auto times = sparsifyme::spmma(A.data().get(), B.data().get(), C.data().get(), m, n, k, (std::size_t)1);
  cudaDeviceSynchronize();
  dump(std::string("spmma_") + tag + "_a_out", A);
  dump(std::string("spmma_") + tag + "_c", C);
  write_meta(std::string("spmma_") + tag + " " + std::to_string(m) + " " + std::to_string(n) + " " + std::to_string(k) +
             " times " + std::to_string(times.size()));
}

// ---- batched::spmm over ell_t (examples/spmm.cu:40-116: block 2, ell_cols = k/2, ascending unique ids) ----
void case_spmm(std::size_t m, std::size_t n, std::size_t k, std::size_t batch) {
  using namespace sparsifyme;
  thrust::host_vector<ell_t<float, memory_space_t::host>> h_As(batch);
  thrust::host_vector<ell_t<float, memory_space_t::device>> d_As(batch);
  for (std::size_t b = 0; b < batch; ++b) {
    auto& A = h_As[b];
    A.rows = m;
    A.cols = k;
    A.block_size = 2;
    A.ell_cols = k / 2;
    A.blocked_rows = util::ceil_div(A.rows, A.block_size);
    A.blocked_cols = util::ceil_div(A.ell_cols, A.block_size);
    A.num_blocks = A.blocked_rows * A.blocked_cols;
    A.column_indices.resize(A.num_blocks);
    A.values.resize(A.ell_cols * A.rows);
    for (std::size_t i = 0; i < A.values.size(); ++i) A.values[i] = uniform(10 + b, i);
    // every block-row keeps every other block column, starting at a pseudo-random phase: ascending and unique
    const std::size_t nbc = A.cols / A.block_size;
    for (std::size_t r = 0; r < A.blocked_rows; ++r) {
      const std::size_t phase = mix(1000 * b + r) & 1;
      for (std::size_t c = 0; c < A.blocked_cols; ++c) {
        std::size_t id = 2 * c + phase;
        if (id >= nbc) id = nbc - 1;
        A.column_indices[A.blocked_cols * r + c] = id;
      }
    }
    d_As[b] = A;
    dump("spmm_ids_" + std::to_string(b), A.column_indices);
    dump("spmm_vals_" + std::to_string(b), A.values);
  }
  thrust::host_vector<float> hB = random_f32(4, k * n);
  thrust::device_vector<float> dB = hB;
  dump("spmm_b", hB);
  thrust::host_vector<float*> C_ptrs;
  thrust::host_vector<thrust::device_vector<float>> dC(batch);
  for (std::size_t b = 0; b < batch; ++b) {
    dC[b].resize(m * n);
    C_ptrs.push_back(dC[b].data().get());
  }
  const float ms = batched::spmm(d_As.data(), dB.data().get(), C_ptrs.data(), m, n, k, batch);
  cudaDeviceSynchronize();
  for (std::size_t b = 0; b < batch; ++b) dump("spmm_c_" + std::to_string(b), dC[b]);
  write_meta("spmm " + std::to_string(m) + " " + std::to_string(n) + " " + std::to_string(k) + " " + std::to_string(batch) +
             " ms " + std::to_string(ms));
}

// ---- batched::strided_coo (examples/batched_coo.cu:102-111, with a well-formed COO) ----
void case_coo(int m, int n, int k, int batch) {
  std::vector<int> rows, cols;
  std::vector<float> vals;
  for (int r = 0; r < m; ++r)
    for (int c = 0; c < k; ++c)
      if ((mix((std::uint64_t)r * 7919u + c) & 7) < 3) {  // ~37 % non-zeros, row-sorted, ascending columns
        rows.push_back(r);
        cols.push_back(c);
        vals.push_back(uniform(5, (std::uint64_t)r * k + c));
      }
  thrust::device_vector<int> d_rows(rows.begin(), rows.end()), d_cols(cols.begin(), cols.end());
  thrust::device_vector<float> d_vals(vals.begin(), vals.end());
  thrust::host_vector<float> hB = random_f32(6, (std::size_t)k * n * batch);
  thrust::device_vector<float> dB = hB, dCv((std::size_t)m * n * batch);
  float* dC = dCv.data().get();
  dump("coo_rows", rows.data(), rows.size());
  dump("coo_cols", cols.data(), cols.size());
  dump("coo_vals", vals.data(), vals.size());
  dump("coo_b", hB);
  const float ms = sparsifyme::batched::strided_coo<float>(m, k, vals.size(), k, n, batch, d_rows.data().get(),
                                                           d_cols.data().get(), d_vals.data().get(), dB.data().get(), &dC,
                                                           1.0f, 0.0f);
  cudaDeviceSynchronize();
  dump("coo_c", dCv);
  write_meta("coo " + std::to_string(m) + " " + std::to_string(n) + " " + std::to_string(k) + " " + std::to_string(batch) +
             " nnz " + std::to_string(vals.size()) + " ms " + std::to_string(ms));
}

// ---- batched::gemm (examples/gemm.cu:60-95: per-batch A, one shared B, device pointer arrays) ----
template <typename T>
void case_gemm(const char* tag, std::size_t m, std::size_t n, std::size_t k, std::size_t batch) {
  thrust::host_vector<T> hA = cast_to<T>(random_f32(7, m * k * batch)), hB = cast_to<T>(random_f32(8, k * n));
  thrust::device_vector<T> dA = hA, dB = hB, dC(m * n * batch);
  thrust::device_vector<T*> pA, pB, pC;
  for (std::size_t b = 0; b < batch; ++b) {
    pA.push_back(dA.data().get() + b * m * k);
    pB.push_back(dB.data().get());
    pC.push_back(dC.data().get() + b * m * n);
  }
  dump(std::string("gemm_") + tag + "_a", hA);
  dump(std::string("gemm_") + tag + "_b", hB);
  const float ms = sparsifyme::batched::gemm(pA.data().get(), pB.data().get(), pC.data().get(), m, n, k, batch);
  cudaDeviceSynchronize();
  dump(std::string("gemm_") + tag + "_c", dC);
  write_meta(std::string("gemm_") + tag + " " + std::to_string(m) + " " + std::to_string(n) + " " + std::to_string(k) + " " +
             std::to_string(batch) + " ms " + std::to_string(ms));
}

}  // namespace

int main(int argc, char** argv) {
  if (argc != 2) {
    std::printf("usage: header_parity <out_dir>\n");
    return 2;
  }
  g_dir = argv[1];
  std::remove((g_dir + "/meta.txt").c_str());
  case_sparsify(36, 52);
  case_spmma<__half>("f16", 128, 256, 192);
  case_spmma<float>("f32", 64, 136, 128);
  case_spmm(72, 40, 64, 3);
  case_coo(96, 56, 160, 2);
  case_gemm<float>("f32", 200, 72, 104, 3);
  case_gemm<__half>("f16", 136, 64, 96, 2);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::printf("ok\n");
  return 0;
}

/*
 * spfy_b200.h -- C ABI of libsparsifyme_b200.so
 *
 * The B200-native (sm_100a) replacement for the device work behind the
 * header-only `include/sparsify.me` API of owensgroup/sparsify.me.  The
 * reference has no FFI: its "plugin boundary" is five function templates
 * (SURVEY.md section 8b).  Every entry point below is what the corresponding
 * template in our `include/sparsify.me/ *.hxx` forwards to, and each one names
 * the reference call it replaces (file:line under the reference tree).
 *
 * Conventions (all entry points):
 *   - plain C types only: raw device pointers, sizes, enums as int;
 *   - return 0 on success, a negative SPFY_E_* code otherwise; never throw;
 *     `spfy_last_error_string()` returns a thread-local description;
 *   - all device work is enqueued on the caller's stream (a cudaStream_t,
 *     spelled `spfy_stream_t` so that this header needs no CUDA include);
 *   - no hidden allocation on the hot path: temporaries come from a
 *     caller-provided workspace, sized by the matching *_workspace_bytes query;
 *   - there is NO host fallback: without a CUDA device every compute entry
 *     point fails with SPFY_E_CUDA.
 */
#ifndef SPFY_B200_H_
#define SPFY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPFY_API
#else
#define SPFY_API __attribute__((visibility("default")))
#endif

typedef struct CUstream_st* spfy_stream_t; /* == cudaStream_t */

/* ---- enums (passed as int) --------------------------------------------- */
enum { SPFY_F16 = 0, SPFY_BF16 = 1, SPFY_F32 = 2, SPFY_F64 = 3 };

/* prune24 modes.  STRIP_MAG is the contract mode (top-2-of-4 by |x| along K,
 * tie -> lower index).  TILE_MAG is the 4x4-tile variant the reference asks
 * cusparseLt for at include/sparsify.me/spmma.hxx:86. */
enum { SPFY_PRUNE_STRIP_MAG = 0, SPFY_PRUNE_TILE_MAG = 1 };

/* layouts of the compressed operand.
 *   CANONICAL : values  [rows][ceil4(cols)/2]   row-major, ascending index
 *               meta    [rows][ceil4(cols)/8]   bytes; group g of a row sits in
 *                       byte g/2, low nibble for even g; nibble = i0 | i1<<2
 *   SM100     : what spfy_spmma consumes.  values and metadata are split into
 *               (128-row x 128-logical-k) tiles; tile (mt,kt) sits at index
 *               kt*m_tiles + mt (k-tile major: the m-tiles of one k-tile are adjacent).
 *               A value tile (16 KiB) is ONE 128-row x 128-byte image, the 128B-swizzled
 *               K-major shared-memory layout the UMMA descriptor reads: row r at r*128,
 *               16-byte chunk c of the row (8 stored values = 16 logical k) at
 *               ((c ^ (r & 7)) << 4).  A metadata tile (2 KiB) is the `tcgen05.cp.128x128b`
 *               source image: the 16-bit word of in-tile row r and 16-column unit q sits at byte
 *               (r>>4)*256 + (q&1)*128 + (r&7)*16 + (q>>1)*4 + ((r>>3)&1)*2.
 *               Padding rows / columns hold +0 and the neutral nibble 0x4.
 *               See DESIGN.md "Data layout in HBM"; pinned by tests/test_boundary.py. */
enum { SPFY_LAYOUT_CANONICAL = 0, SPFY_LAYOUT_SM100 = 1 };

enum { SPFY_OP_N = 0, SPFY_OP_T = 1 }; /* == cusparseOperation_t values */
/* OR-ed into the `opB` argument of spfy_spmma / spfy_spmma_problem: D is written TRANSPOSED, [n][m] row-major with pitch
 * ldd >= m -- for a convolution layer that is NHWC, what spfy_spmma_conv of the next layer reads.  beta must be 0 and
 * m a multiple of 8 (the contiguous dimension of an output always is: n for the row-major one). */
enum { SPFY_OUT_T = 0x10 };

enum {
  SPFY_OK = 0,
  SPFY_E_INVALID = -1,     /* bad argument */
  SPFY_E_UNSUPPORTED = -2, /* valid but not implemented for these arguments */
  SPFY_E_CUDA = -3,        /* CUDA runtime / driver error (see last error)   */
  SPFY_E_WORKSPACE = -4,   /* workspace too small                            */
  SPFY_E_NCCL = -5
};

SPFY_API int spfy_version(void);
/* Loads the library's device code on the current device (CUDA loads kernels lazily at first use, tens of
 * milliseconds for the large ones).  Optional and idempotent; the header templates call it before they start
 * their timers -- the counterpart of the handle / plan creation the reference keeps outside its timers
 * (spmma.hxx:51-80). */
SPFY_API int spfy_init(void);
SPFY_API const char* spfy_last_error_string(void);
/* number of kernels this library launched in this process (bench: gpu_launches) */
SPFY_API uint64_t spfy_launch_count(void);

/* Element-wise dtype conversion between F32 and F16/BF16 (round to nearest even), used by
 * the header templates when a driver instantiates spmma<float> (examples/spmma.cu:24). */
SPFY_API int spfy_convert(int src_dtype, int dst_dtype, const void* src, void* dst, size_t count,
                          spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * A1  sparsifyme::sparsify<BLK_M,BLK_N>          include/sparsify.me/sparsify.hxx:24-82
 * Exact positional semantics of the reference: mask <- 1 (:71); then for each
 * of (m/blk_m)*(n/blk_n) linear blocks zero the first
 * floor(blk_m*blk_n*sparsity_factor) offsets of the sequence h + w*blk_n,
 * h outer / w inner (:53-65), in `weights` and in `mask`.  One fused kernel
 * replaces thrust::fill_n + thrust::transform.  Writes that the reference
 * would issue past m*n (BLK_N > BLK_M instantiations) are dropped.
 * `mask` is an array of m*n 64-bit words (std::size_t in the reference).
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_prune_blocks_ref(int dtype, void* weights, uint64_t* mask,
                                   size_t m, size_t n, size_t blk_m, size_t blk_n,
                                   float sparsity_factor, spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * A2+A3  cusparseLtSpMMAPrune / PruneCheck / CompressedSize / Compress
 *                                               include/sparsify.me/spmma.hxx:85-104
 * One bandwidth-bound kernel: reads `in` (rows x cols, row-major, leading
 * dimension ld_in elements) once and writes any subset of
 *   out_dense : the pruned dense matrix (may alias `in`: in-place like :86)
 *   comp_vals / meta : compressed operand in `layout`
 *   mask      : rows*cols 64-bit keep flags (1 = kept), API parity with A1
 * Null outputs are skipped.  dtype is SPFY_F16 or SPFY_BF16.
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_compressed_bytes(int dtype, size_t rows, size_t cols, int layout,
                                   size_t* vals_bytes, size_t* meta_bytes);
SPFY_API int spfy_prune24(int dtype, int mode, int layout, const void* in, size_t ld_in,
                          void* out_dense, size_t ld_out, void* comp_vals, void* meta,
                          uint64_t* mask, size_t rows, size_t cols, spfy_stream_t stream);
/* Whole-model variant: prune+compress `count` matrices in as few launches as possible (one
 * per 96 matrices).  The per-layer weight matrices of datasets/ *.csv are <= 4.7 MB, so one
 * launch per layer is launch-latency-bound; this is the call the prune GB/s figure is quoted
 * on (mode STRIP_MAG).  TILE_MAG -- the algorithm the reference requests (spmma.hxx:86) --
 * needs out_dense per item (it may alias the input) and takes the one-pass prune + compress
 * kernel for every matrix with cols % 16 == 0 and 8-byte aligned rows; the others (k = 147)
 * go through spfy_prune24 one by one.  `items` is a HOST array; fields as in spfy_prune24
 * (null outputs are skipped). */
typedef struct spfy_prune24_item {
  const void* in;
  size_t ld_in;
  void* out_dense;
  size_t ld_out;
  void* comp_vals;
  void* meta;
  size_t rows, cols;
} spfy_prune24_item;
SPFY_API int spfy_prune24_batched(int dtype, int mode, int layout, const spfy_prune24_item* items,
                                  size_t count, spfy_stream_t stream);
/* A dense matrix obeys 2:4 along its rows iff *d_invalid == 0 afterwards
 * (same convention as cusparseLtSpMMAPruneCheck, spmma.hxx:88-94). */
SPFY_API int spfy_prune24_check(int dtype, const void* in, size_t ld_in, size_t rows,
                                size_t cols, int* d_invalid, spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * A4  cusparseLtMatmul                           include/sparsify.me/spmma.hxx:106-114
 * D = alpha * A(2:4) * op(B) + beta * C, fp32 accumulate on tcgen05.mma.sp.
 * A is the SM100-layout compressed operand of an m x k matrix (from
 * spfy_prune24).  All dense operands are row-major like the reference's
 * descriptors (:56-64): op(B) is k x n.  opB = N: B is k x n, ldb >= n;
 * opB = T: B is n x k, ldb >= k.  C and D are m x n; D may alias C.
 * n, ldb, ldc, ldd must be multiples of 8 elements and all base pointers
 * 16-byte aligned (the reference's own fp16 contract, spmma.hxx:45-49).
 * opB | SPFY_OUT_T: D is n x m (the transpose, pitch ldd >= m), beta must be 0, m % 8 == 0.
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_spmma_workspace_bytes(int dtype, size_t m, size_t n, size_t k,
                                        size_t* bytes);
SPFY_API int spfy_spmma(int dtype, int opB, size_t m, size_t n, size_t k, float alpha,
                        const void* comp_vals, const void* meta, const void* B, size_t ldb,
                        float beta, const void* C, size_t ldc, void* D, size_t ldd,
                        void* workspace, size_t workspace_bytes, spfy_stream_t stream);

/* Implicit GEMM: the step BEFORE the path fused into it.  The K x N operand of every CSV row
 * is the `unfold` of a convolution input (datasets/get_shapes.py:29-41 of the reference): a 3 x 3 layer reads each
 * activation nine times.  Here B is never materialised: X is the NHWC activation tensor [batch][h][w][c] and the
 * kernel's producer gathers each (128 positions x 64 channels) piece of a B stage with one TMA im2col instruction.
 *     D[m x N] = alpha * A(2:4)[m x K] * im2col(X)[K x N] + beta * C,   N = batch * ho * wo  (image, row, column),
 *     K = kh * kw * c ordered (kh, kw, c): permute the weight columns from PyTorch's (c, kh, kw) BEFORE pruning
 *     (spfy_permute_conv_weights), the 2:4 groups are formed along the new order.
 * c must be a multiple of 64 (every ResNet layer but the first; that one is unfolded explicitly), dilation 1.
 * D is row-major m x N like spfy_spmma (ldd % 8 == 0). */
typedef struct spfy_conv_desc {
  size_t batch, h, w, c; /* NHWC input */
  size_t kh, kw;         /* filter */
  size_t stride, pad;    /* same in both directions */
} spfy_conv_desc;
SPFY_API int spfy_spmma_conv(int dtype, const spfy_conv_desc* conv, size_t m, float alpha,
                             const void* comp_vals, const void* meta, const void* X, float beta,
                             const void* C, size_t ldc, void* D, size_t ldd, spfy_stream_t stream);
/* The same with the output in NHWC: Y[N][m] (pitch ldy >= m, ldy % 8 == 0), i.e. [batch][ho][wo][m] -- the X of the next
 * layer's spfy_spmma_conv, so a chain of convolutions never transposes or unfolds anything.  A unit's 128 positions x m
 * channels are then one contiguous run of Y instead of m rows 2 * N bytes apart.  beta = 0. */
SPFY_API int spfy_spmma_conv_nhwc(int dtype, const spfy_conv_desc* conv, size_t m, float alpha,
                                  const void* comp_vals, const void* meta, const void* X, void* Y, size_t ldy,
                                  spfy_stream_t stream);
/* weights [m][c * kh * kw] in (c, kh, kw) column order (what torch's unfold / a flattened conv weight uses)
 * -> [m][kh * kw * c] in (kh, kw, c) order; 16-bit elements */
SPFY_API int spfy_permute_conv_weights(const void* in, void* out, size_t m, size_t c, size_t kh, size_t kw,
                                       spfy_stream_t stream);

/* Many independent problems (the per-layer GEMMs of a datasets/ *.csv table) as ONE plan:
 * tensor maps and the tile schedule are built once (like cusparseLtMatmulPlanInit,
 * spmma.hxx:79, which the reference also keeps outside its timers) and every run issues at
 * most one persistent launch per ring-geometry class (six) and operand orientation that walk all problems' tiles, so neither launch latency
 * nor per-layer wave quantisation is paid per layer.  Problems must not alias each other's
 * outputs.  plan_create allocates a small device table; plan_run never allocates or syncs. */
typedef struct spfy_spmma_problem {
  int opB;                 /* SPFY_OP_N / SPFY_OP_T */
  size_t m, n, k;
  const void* comp_vals;   /* SM100-layout compressed A (spfy_prune24) */
  const void* meta;
  const void* B;
  size_t ldb;
  const void* C;           /* may be null when beta == 0 */
  size_t ldc;
  void* D;
  size_t ldd;
  float alpha, beta;
} spfy_spmma_problem;
typedef struct spfy_spmma_plan_st* spfy_spmma_plan_t;
SPFY_API int spfy_spmma_plan_create(int dtype, const spfy_spmma_problem* problems, size_t count,
                                    spfy_spmma_plan_t* plan);
/* Replicated outputs -- the fused form of the output gather (north_star: the path's only exchange).  Every D tile of
 * problem i is stored, by the same TMA store that writes problems[i].D, to `replicas` further matrices
 * replica_D[i * replicas + r] of the same shape and ldd.  With one process per GPU these are this rank's slab of the
 * gather arena in every PEER GPU's memory (spfy_peer_open mappings, reached over NVLink): the GEMM epilogue is then
 * the all-gather, tile by tile, and no collective follows -- only a barrier among the ranks before anyone reads
 * (completion of the launch on rank r makes rank r's stores visible; it says nothing about the other ranks').
 * replicas <= 15; replica pointers 16-byte aligned; beta != 0 still reads problems[i].C only. */
SPFY_API int spfy_spmma_plan_create_replicated(int dtype, const spfy_spmma_problem* problems, size_t count,
                                               size_t replicas, void* const* replica_D, spfy_spmma_plan_t* plan);
/* A plan whose problems may be convolution layers (a whole network's table with the 3 x 3 layers never unfolded):
 * convs[i] != NULL makes problem i the implicit GEMM of spfy_spmma_conv -- problems[i].B is then the NHWC activation
 * tensor X, n / k follow from the descriptor (problems[i].n, .k, .ldb and the transpose bits of .opB are ignored;
 * SPFY_OUT_T in .opB still selects the NHWC output of spfy_spmma_conv_nhwc), comp_vals / meta hold the weights with K in
 * (kh, kw, c) order.  convs == NULL or convs[i] == NULL: as spfy_spmma_plan_create.  Descriptors are read at creation
 * only.  Results are bitwise those of the single calls. */
SPFY_API int spfy_spmma_plan_create_conv(int dtype, const spfy_spmma_problem* problems,
                                         const spfy_conv_desc* const* convs, size_t count, spfy_spmma_plan_t* plan);
SPFY_API int spfy_spmma_plan_run(spfy_spmma_plan_t plan, spfy_stream_t stream);
SPFY_API int spfy_spmma_plan_launches(spfy_spmma_plan_t plan); /* kernel launches per run */
/* introspection / profiling: run one of the plan's launches, or describe it */
SPFY_API int spfy_spmma_plan_run_launch(spfy_spmma_plan_t plan, int index, spfy_stream_t stream);
SPFY_API int spfy_spmma_plan_launch_info(spfy_spmma_plan_t plan, int index, int* problems, int* units,
                                         int* stages, int* smem_bytes);
SPFY_API int spfy_spmma_plan_destroy(spfy_spmma_plan_t plan);

/* ------------------------------------------------------------------------
 * Pruned-layer container (the artifact that follows prune + compress).
 * One self-describing HOST buffer per compressed 2:4 operand:
 *   64-byte header {magic "SPFY24\0\0", version, dtype, layout, rows, cols, vals_bytes,
 *   meta_bytes, FNV-1a 64 of the payload} + values + metadata, as spfy_prune24 wrote them
 * (copy them to the host first; these calls touch host memory only, no CUDA).  A reader
 * checks magic / version / sizes against spfy_compressed_bytes and the checksum, then
 * hands back offsets into the buffer, so the payload can go to the device with two copies
 * and be used by spfy_spmma as is.
 * ---------------------------------------------------------------------- */
#define SPFY_PACKED_HEADER_BYTES 64
SPFY_API int spfy_packed_bytes(int dtype, size_t rows, size_t cols, int layout, size_t* bytes);
SPFY_API int spfy_packed_write(int dtype, int layout, size_t rows, size_t cols,
                               const void* host_vals, const void* host_meta, void* dst,
                               size_t dst_bytes);
SPFY_API int spfy_packed_read(const void* src, size_t src_bytes, int* dtype, int* layout,
                              size_t* rows, size_t* cols, size_t* vals_offset,
                              size_t* vals_bytes, size_t* meta_offset, size_t* meta_bytes);

/* ------------------------------------------------------------------------
 * Unstructured path (north_star subsystem 3).  Threshold prune keeps x iff
 * |x| > threshold (compared in fp32) and emits COO sorted by (row, col) or CSR.
 * dtype of `in` is F16/BF16/F32; emitted values are fp32 like the reference's
 * CUDA_R_32F COO (include/sparsify.me/spmm.hxx:165-168).
 * nnz is produced on the device (`d_nnz`, one int64) so the call never syncs.
 * `capacity` bounds the number of entries written.  `d_row_ptr_or_null`
 * (rows+1 int32), when given, receives the CSR row pointer of the same entries.
 * One pass over the input: one memset of the workspace (8 bytes per 4096
 * elements, size from the query; must be 8-byte aligned) + one kernel.
 * rows * cols must be below 2^31.
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_threshold_workspace_bytes(size_t rows, size_t cols, size_t* bytes);
SPFY_API int spfy_threshold_to_coo(int dtype, const void* in, size_t ld_in, size_t rows,
                                   size_t cols, float threshold, int32_t* row_idx,
                                   int32_t* col_idx, float* vals, size_t capacity,
                                   int64_t* d_nnz, int32_t* d_row_ptr_or_null,
                                   void* workspace, size_t workspace_bytes,
                                   spfy_stream_t stream);
/* COO (sorted by row) -> CSR row pointer (rows+1 int32). */
SPFY_API int spfy_coo_to_csr(const int32_t* row_idx, size_t nnz, size_t rows,
                             int32_t* row_ptr, spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * Algorithm of the unstructured SpMM entry points below (first argument).
 *   DEFAULT     : the library chooses.  An A with at least 2 % non-zeros (every pruned ResNet
 *                 weight matrix: 5-50 %) is scattered into a dense matrix in the workspace and
 *                 contracted on tcgen05 (fp32 operands: 3xTF32, fp32-level accuracy, see
 *                 spfy_gemm_*; 16-bit operands: kind::f16) when the operands meet the TMA
 *                 contract (16-byte aligned bases, leading dimensions and batch strides
 *                 multiples of 16 bytes); otherwise, and for sparser A, the CUDA-core kernels run.
 *                 The CSR entry does not know nnz on the host: it launches both candidates and
 *                 a device flag lets exactly one do the work, so it never reads anything back.
 *   CUDA_CORE   : the row-split CUDA-core kernels only.  fp32 sums in ascending k order per
 *                 output element: bit-identical to cuSPARSE 12.5 on exactly representable
 *                 inputs (tests/golden/cusparse_*.npz).
 *   TENSOR      : the tensor-core route or SPFY_E_UNSUPPORTED / SPFY_E_WORKSPACE.
 *   TENSOR_FAST : same with ONE TF32 product per fp32 multiply (10-bit mantissas; the accuracy
 *                 class of cuBLAS's TF32 mode).  16-bit operands: identical to TENSOR.
 * ---------------------------------------------------------------------- */
enum {
  SPFY_SPMM_ALG_DEFAULT = 0,
  SPFY_SPMM_ALG_CUDA_CORE = 1,
  SPFY_SPMM_ALG_TENSOR = 2,
  SPFY_SPMM_ALG_TENSOR_FAST = 3
};

/* ------------------------------------------------------------------------
 * A6  sparsifyme::batched::strided_coo           include/sparsify.me/spmm.hxx:140-193
 * C_b = alpha * A * B_b + beta * C_b for b in [0, num_batches): ONE sparse A
 * (stride 0, :169) shared by all batches; B_b = B + b*strideB is k x n
 * column-major (ldb >= k, :160,:170); C_b = C + b*strideC is m x n column-major
 * (ldc >= m, :161,:173).  fp32 values, int32 indices, fp32 accumulate.
 * The COO triplets must be sorted by row (any column order; repeated (row, col)
 * entries add, like cuSPARSE); row_ptr is built internally in the workspace.
 * The workspace query takes the algorithm and the batch extent: the tensor-core route keeps
 * the dense m x k fp32 image of A there and, when ldb = k is not a multiple of 16 bytes
 * (k = 147), a padded copy of every B_b (a workspace sized for CUDA_CORE makes DEFAULT
 * take the CUDA-core kernels).
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_spmm_workspace_bytes(int alg, size_t m, size_t k, size_t n, size_t num_batches,
                                       size_t nnz, size_t* bytes);
SPFY_API int spfy_spmm_coo_strided_batched(int alg, size_t m, size_t k, size_t nnz, size_t n,
                                           size_t num_batches, const int32_t* row_idx,
                                           const int32_t* col_idx, const float* vals,
                                           const float* B, size_t ldb, size_t strideB,
                                           float* C, size_t ldc, size_t strideC, float alpha,
                                           float beta, void* workspace, size_t workspace_bytes,
                                           spfy_stream_t stream);
/* Same product with A already in CSR (same workspace query). */
SPFY_API int spfy_spmm_csr_strided_batched(int alg, size_t m, size_t k, size_t n,
                                           size_t num_batches, const int32_t* row_ptr,
                                           const int32_t* col_idx, const float* vals,
                                           const float* B, size_t ldb, size_t strideB, float* C,
                                           size_t ldc, size_t strideC, float alpha, float beta,
                                           void* workspace, size_t workspace_bytes,
                                           spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * A5  sparsifyme::batched::spmm (blocked-ELL)    include/sparsify.me/spmm.hxx:30-138
 * For each batch b: C_b = alpha * A_b * B + beta * C_b.  A_b is blocked-ELL
 * (containers/ell.hxx:24-33): rows x cols, square blocks of `block`,
 * `ell_cols` stored columns per row; col_idx_b[(rows/block) x (ell_cols/block)]
 * block-column ids (int64 here because the reference stores std::size_t,
 * ell.hxx:31; ids < 0 are padding); values_b[rows x ell_cols] row-major.  B is
 * k x n column-major ldb = k shared by all batches (:67); C_b is m x n
 * column-major ldc = m (:63).  `col_idx`, `values`, `Cs` are DEVICE arrays of
 * num_batches device pointers.  dtype is that of values/B/C (F16/BF16/F32);
 * fp32 accumulate (:82).
 * Tensor-core route (DEFAULT when the operands meet the TMA contract): the batch is
 * expanded chunk by chunk into dense row-major matrices in the workspace (every
 * element written once, no atomics) and contracted by the dense tcgen05 GEMM; a
 * block-row that repeats an id (undefined for cuSPARSE; our kernels add the blocks)
 * raises a device flag that hands its chunk to the CUDA-core kernel instead.
 * ---------------------------------------------------------------------- */
SPFY_API int spfy_spmm_bell_workspace_bytes(int alg, int dtype, size_t rows, size_t cols, size_t n,
                                            size_t num_batches, size_t* bytes);
SPFY_API int spfy_spmm_bell_batched(int alg, int dtype, size_t rows, size_t cols, size_t n,
                                    size_t block, size_t ell_cols, size_t num_batches,
                                    const int64_t* const* col_idx, const void* const* values,
                                    const void* B, size_t ldb, void* const* Cs, size_t ldc,
                                    float alpha, float beta, void* workspace,
                                    size_t workspace_bytes, spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * N3  sparsifyme::batched::gemm                  include/sparsify.me/gemm.hxx:25-195
 * Dense batched GEMM on tcgen05 (replaces cublas{H,S}gemmBatched, :51-186), BLAS
 * conventions: column-major, C_b[m x n, ldc] = alpha * op(A_b)[m x k] * op(B_b)[k x n]
 * + beta * C_b.  opX = N: the operand is stored as its op() shape; T: transposed.
 * dtype F16 / BF16: kind::f16, fp32 accumulation in TMEM.  dtype F32:
 *   SPFY_GEMM_PRECISE  3xTF32 -- each operand is split in shared memory into
 *                      hi (the 11 significant bits the tensor core reads) and lo = x - hi, and
 *                      hi*lo + lo*hi + hi*hi is accumulated in fp32: products exact to ~2^-21;
 *   SPFY_GEMM_FAST     one TF32 product.
 * SPFY_GEMM_CTA_PAIRS may be OR-ed into `precision`: 16-bit operands and SPFY_GEMM_FAST then run on CTA pairs
 * (tcgen05.mma.cta_group::2, M = 256: each CTA of a two-CTA cluster loads its own 128 rows of the long operand and
 * half of the other tile) when every problem has more than 128 rows of the long operand; same results.  Off by
 * default because it is slower on the ResNet shapes (profiles/r02_gemm_cta_pairs.txt).
 * The operands are fetched by TMA, which needs 16-byte aligned bases and lda / ldb /
 * batch strides that are multiples of 16 bytes; an operand that misses this (ldb = k = 147
 * floats, the first conv layer of every ResNet) is first copied into the workspace with
 * padded rows -- one extra pass over that operand, nothing else changes.  The batched form
 * takes HOST arrays of device pointers (what examples/gemm.cu:93-95 holds are device
 * arrays -- the header copies them back once, outside its timer).  Both forms take a
 * device workspace sized by the query (problem table of the batched form + padded copies
 * of the operands whose lda / ldb need them; pass lda = 0 / ldb = 0 to size for a
 * misaligned base pointer as well).
 * ---------------------------------------------------------------------- */
enum { SPFY_GEMM_PRECISE = 0, SPFY_GEMM_FAST = 1, SPFY_GEMM_CTA_PAIRS = 0x10 };
SPFY_API int spfy_gemm_workspace_bytes(int dtype, int opA, int opB, size_t m, size_t n, size_t k,
                                       size_t lda, size_t ldb, size_t num_batches, size_t* bytes);
SPFY_API int spfy_gemm_strided_batched(int dtype, int precision, int opA, int opB, size_t m, size_t n,
                                       size_t k, float alpha, const void* A, size_t lda,
                                       size_t strideA, const void* B, size_t ldb, size_t strideB,
                                       float beta, void* C, size_t ldc, size_t strideC,
                                       size_t num_batches, void* workspace, size_t workspace_bytes,
                                       spfy_stream_t stream);
SPFY_API int spfy_gemm_batched(int dtype, int precision, int opA, int opB, size_t m, size_t n, size_t k,
                               float alpha, const void* const* A_ptrs, size_t lda,
                               const void* const* B_ptrs, size_t ldb, float beta,
                               void* const* C_ptrs, size_t ldc, size_t num_batches,
                               void* workspace, size_t workspace_bytes, spfy_stream_t stream);

/* ------------------------------------------------------------------------
 * Multi-GPU: the per-layer problems and, inside one GEMM, the N = batch x
 * spatial columns are independent, so the path shards with NO exchange during compute;
 * the only collective is a gather of outputs (north_star).  The reference has no
 * multi-GPU code (examples/spmma.cu:27-28 queries device 0 only).  One process per GPU.
 * NCCL is loaded at run time (libnccl.so.2, or the path in SPFY_NCCL_LIB): without it these
 * calls return SPFY_E_NCCL and nothing else in the library is affected.
 *   bootstrap : rank 0 calls spfy_mg_unique_id (128 bytes), the host program hands the
 *               bytes to every rank (torch.distributed, MPI, a file ...), all call
 *               spfy_mg_create with the current device set.
 *   allgather : N-sharded GEMMs.  Rank r holds D_r [M x N/g] row-major (its images' columns);
 *               the gathered result is [g][M][N/g] -- the ranks' slabs back to back.  In
 *               place when `send` == `recv` + r * bytes_per_rank: let the GEMM write D_r
 *               there and no copy is made at all.
 *   broadcast_many : layer sharding.  Entry i is sent from rank roots[i] to everybody, all
 *               entries as ONE NCCL group (the lists are HOST arrays).
 * All work is enqueued on the caller's stream.
 * ---------------------------------------------------------------------- */
/* Peer memory for the fused gather (spfy_spmma_plan_create_replicated): device memory another process on the same box
 * can map.  spfy_peer_alloc = cudaMalloc + cudaIpcGetMemHandle (the 64 handle bytes travel through the host program,
 * like the NCCL id); spfy_peer_open = cudaIpcOpenMemHandle with lazy peer access, in a DIFFERENT process than the
 * owner; the mapping is an ordinary device pointer for every entry point of this library. */
#define SPFY_PEER_HANDLE_BYTES 64
SPFY_API int spfy_peer_alloc(size_t bytes, void** ptr, void* handle64);
SPFY_API int spfy_peer_free(void* ptr);
SPFY_API int spfy_peer_open(const void* handle64, void** ptr);
SPFY_API int spfy_peer_close(void* ptr);

typedef struct spfy_mg_comm_st* spfy_mg_comm_t;
SPFY_API int spfy_mg_unique_id(void* id128);
SPFY_API int spfy_mg_create(int rank, int world, const void* id128, spfy_mg_comm_t* comm);
SPFY_API int spfy_mg_destroy(spfy_mg_comm_t comm);
SPFY_API int spfy_mg_rank(spfy_mg_comm_t comm);
SPFY_API int spfy_mg_world(spfy_mg_comm_t comm);
SPFY_API int spfy_mg_allgather(spfy_mg_comm_t comm, const void* send, void* recv, size_t bytes_per_rank,
                               spfy_stream_t stream);
SPFY_API int spfy_mg_broadcast_many(spfy_mg_comm_t comm, void* const* buffers, const size_t* bytes,
                                    const int* roots, size_t count, spfy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPFY_B200_H_ */

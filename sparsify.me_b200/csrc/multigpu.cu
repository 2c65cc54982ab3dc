// multigpu.cu -- the only collectives of the path: gathering outputs across the GPUs of one box (north_star:
// "NCCL used only to gather outputs").  One process per GPU; the reference has no multi-GPU
// code at all (examples/spmma.cu:27-28 queries device 0 only).
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 -- the copy PyTorch ships, or SPFY_NCCL_LIB), so the library
// has no link-time dependency on it and a single-GPU user never touches it.  Bootstrap is NCCL's own: rank 0 makes a
// unique id (spfy_mg_unique_id), the host program distributes its 128 bytes however it likes (torch.distributed,
// MPI, a file), every rank calls spfy_mg_create.
//
// Layouts (documented in include/spfy_b200.h): an N-sharded GEMM leaves rank r with D_r [M x N/g] row-major; the
// gathered result is [g][M][N/g] -- the ranks' slabs back to back, which is exactly ncclAllGather, in place when
// D_r already sits at slab r of the receive buffer: no transpose, no concatenation pass.
#include "common.cuh"

#include <dlfcn.h>

#include <mutex>

namespace spfy {
namespace {

// the handful of NCCL entry points used, by their C signatures (nccl.h: ncclResult_t is an int enum, ncclComm_t an
// opaque pointer, ncclUniqueId 128 bytes, ncclDataType_t ncclUint8 == 1)
struct IdBytes { char b[128]; };  // ncclUniqueId, passed BY VALUE to ncclCommInitRank

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, IdBytes, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.lib) return SPFY_OK;
  const char* names[] = {getenv("SPFY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(SPFY_E_NCCL, "NCCL is not available: dlopen(libnccl.so.2) failed (%s); set SPFY_NCCL_LIB", dlerror());
  NcclApi api;
  api.lib = lib;
#define SPFY_NCCL_SYM(field, name)                                                        \
  *(void**)(&api.field) = dlsym(lib, name);                                               \
  if (!api.field) {                                                                        \
    dlclose(lib);                                                                          \
    return fail(SPFY_E_NCCL, "NCCL library has no symbol %s", name);                       \
  }
  SPFY_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  SPFY_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  SPFY_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  SPFY_NCCL_SYM(AllGather, "ncclAllGather")
  SPFY_NCCL_SYM(Broadcast, "ncclBroadcast")
  SPFY_NCCL_SYM(GroupStart, "ncclGroupStart")
  SPFY_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  SPFY_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SPFY_NCCL_SYM
  g_nccl = api;
  return SPFY_OK;
}

int nccl_fail(const char* what, int rc) {
  return fail(SPFY_E_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
}

struct MgComm {
  void* comm = nullptr;
  int rank = 0, world = 1;
};

}  // namespace
}  // namespace spfy

using namespace spfy;

extern "C" {

int spfy_mg_unique_id(void* id128) {
  if (!id128) return fail(SPFY_E_INVALID, "mg_unique_id: null pointer");
  int rc = load_nccl();
  if (rc) return rc;
  const int n = g_nccl.GetUniqueId(id128);
  return n ? nccl_fail("ncclGetUniqueId", n) : SPFY_OK;
}

int spfy_mg_create(int rank, int world, const void* id128, spfy_mg_comm_t* out) {
  if (!out) return fail(SPFY_E_INVALID, "mg_create: null handle pointer");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(SPFY_E_INVALID, "mg_create: rank %d of %d", rank, world);
  int rc = load_nccl();
  if (rc) return rc;
  MgComm* c = new MgComm();
  c->rank = rank;
  c->world = world;
  IdBytes id;
  memcpy(id.b, id128, sizeof(id.b));
  const int n = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (n) {
    delete c;
    return nccl_fail("ncclCommInitRank", n);
  }
  *out = reinterpret_cast<spfy_mg_comm_t>(c);
  return SPFY_OK;
}

int spfy_mg_destroy(spfy_mg_comm_t h) {
  if (!h) return SPFY_OK;
  MgComm* c = reinterpret_cast<MgComm*>(h);
  const int n = c->comm ? g_nccl.CommDestroy(c->comm) : 0;
  delete c;
  return n ? nccl_fail("ncclCommDestroy", n) : SPFY_OK;
}

int spfy_mg_allgather(spfy_mg_comm_t h, const void* send, void* recv, size_t bytes_per_rank, spfy_stream_t stream) {
  if (!h) return fail(SPFY_E_INVALID, "mg_allgather: null communicator");
  MgComm* c = reinterpret_cast<MgComm*>(h);
  if (bytes_per_rank == 0) return SPFY_OK;
  if (!send || !recv) return fail(SPFY_E_INVALID, "mg_allgather: null buffer");
  const int n = g_nccl.AllGather(send, recv, bytes_per_rank, /*ncclUint8*/ 1, c->comm, (cudaStream_t)stream);
  return n ? nccl_fail("ncclAllGather", n) : SPFY_OK;
}

int spfy_mg_broadcast_many(spfy_mg_comm_t h, void* const* buffers, const size_t* bytes, const int* roots, size_t count,
                           spfy_stream_t stream) {
  if (!h) return fail(SPFY_E_INVALID, "mg_broadcast_many: null communicator");
  MgComm* c = reinterpret_cast<MgComm*>(h);
  if (count == 0) return SPFY_OK;
  if (!buffers || !bytes || !roots) return fail(SPFY_E_INVALID, "mg_broadcast_many: null list");
  int n = g_nccl.GroupStart();
  if (n) return nccl_fail("ncclGroupStart", n);
  for (size_t i = 0; i < count && !n; ++i) {
    if (!bytes[i]) continue;
    if (roots[i] < 0 || roots[i] >= c->world || !buffers[i]) {
      g_nccl.GroupEnd();
      return fail(SPFY_E_INVALID, "mg_broadcast_many: entry %zu: root %d, buffer %p", i, roots[i], buffers[i]);
    }
    n = g_nccl.Broadcast(buffers[i], buffers[i], bytes[i], 1, roots[i], c->comm, (cudaStream_t)stream);
  }
  const int e = g_nccl.GroupEnd();
  if (n) return nccl_fail("ncclBroadcast", n);
  return e ? nccl_fail("ncclGroupEnd", e) : SPFY_OK;
}

// ---- peer memory (fused gather: the spmma epilogue stores into every peer's arena, spfy_spmma_plan_create_replicated)
int spfy_peer_alloc(size_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64) return fail(SPFY_E_INVALID, "peer_alloc: null pointer");
  *ptr = nullptr;
  if (!bytes) return fail(SPFY_E_INVALID, "peer_alloc: zero bytes");
  static_assert(sizeof(cudaIpcMemHandle_t) == SPFY_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(SPFY_E_CUDA, "peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(SPFY_E_CUDA, "peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return SPFY_OK;
}

int spfy_peer_free(void* ptr) {
  if (!ptr) return SPFY_OK;
  const cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? SPFY_OK : fail(SPFY_E_CUDA, "peer_free: %s", cudaGetErrorString(e));
}

int spfy_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(SPFY_E_INVALID, "peer_open: null pointer");
  *ptr = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? SPFY_OK
                          : fail(SPFY_E_CUDA, "peer_open: cudaIpcOpenMemHandle: %s (the owner must be another process "
                                              "on the same box, its GPU reachable by peer access)", cudaGetErrorString(e));
}

int spfy_peer_close(void* ptr) {
  if (!ptr) return SPFY_OK;
  const cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? SPFY_OK : fail(SPFY_E_CUDA, "peer_close: %s", cudaGetErrorString(e));
}

int spfy_mg_rank(spfy_mg_comm_t h) { return h ? reinterpret_cast<MgComm*>(h)->rank : -1; }
int spfy_mg_world(spfy_mg_comm_t h) { return h ? reinterpret_cast<MgComm*>(h)->world : 0; }

}  // extern "C"

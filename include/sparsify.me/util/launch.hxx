// util/launch.hxx -- per-batch launch bundle (reference: include/sparsify.me/util/launch.hxx:19-42).
// The reference pairs every batch element with a private stream AND a cuSPARSE handle; there is
// no cuSPARSE in this build, so the bundle keeps the stream, an event and an optional
// stream-ordered scratch buffer only.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

namespace sparsifyme {
namespace util {

struct launch_t {
  cudaStream_t stream = nullptr;
  cudaEvent_t event = nullptr;
  void* buffer = nullptr;
  std::size_t buffer_size = 0;
};

inline void create_launch_configs(std::vector<launch_t>& configs) {
  for (launch_t& c : configs) {
    cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c.event, cudaEventDisableTiming);
  }
}

inline void destroy_launch_configs(std::vector<launch_t>& configs) {
  for (launch_t& c : configs) {
    if (c.buffer) cudaFreeAsync(c.buffer, c.stream);
    c.buffer = nullptr;
    c.buffer_size = 0;
    if (c.event) cudaEventDestroy(c.event);
    if (c.stream) {
      cudaStreamSynchronize(c.stream);
      cudaStreamDestroy(c.stream);
    }
    c.event = nullptr;
    c.stream = nullptr;
  }
}

}  // namespace util
}  // namespace sparsifyme

// util/launch.hxx -- one launch bundle per batch element (reference: include/sparsify.me/util/launch.hxx:19-42,
// where every element of a batch gets a private stream and a private cuSPARSE handle).  There is no
// cuSPARSE in this build: a bundle is a non-blocking stream, an event to join on, and an optional
// stream-ordered scratch buffer.  The free functions are `inline` (the reference's are not, which
// limits it to one translation unit).
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

  void open() {
    if (!stream) cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (!event) cudaEventCreateWithFlags(&event, cudaEventDisableTiming);
  }
  void close() {
    if (stream) cudaStreamSynchronize(stream);
    if (buffer) cudaFree(buffer);
    if (event) cudaEventDestroy(event);
    if (stream) cudaStreamDestroy(stream);
    *this = launch_t{};
  }
};

inline void create_launch_configs(std::vector<launch_t>& configs) {
  for (std::size_t i = 0; i < configs.size(); ++i) configs[i].open();
}

inline void destroy_launch_configs(std::vector<launch_t>& configs) {
  for (std::size_t i = configs.size(); i-- > 0;) configs[i].close();
}

}  // namespace util
}  // namespace sparsifyme

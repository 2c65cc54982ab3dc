// util/timer.hxx -- cudaEvent stopwatch with the interface the drivers use
// (reference: include/sparsify.me/util/timer.hxx:24-55): begin()/end() record on a stream
// and block the host until the event has happened; end() returns milliseconds.
#pragma once
#include <cuda_runtime.h>

namespace sparsifyme {
namespace util {

class timer_t {
 public:
  float time = 0.f;  // last measured interval, ms (public like the reference's member)

  timer_t() {
    cudaEventCreate(&t0_);
    cudaEventCreate(&t1_);
  }
  ~timer_t() {
    cudaEventDestroy(t0_);
    cudaEventDestroy(t1_);
  }
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;

  void begin(cudaStream_t stream = 0) {
    cudaEventRecord(t0_, stream);
    cudaEventSynchronize(t0_);
  }
  float end(cudaStream_t stream = 0) {
    cudaEventRecord(t1_, stream);
    cudaEventSynchronize(t1_);
    cudaEventElapsedTime(&time, t0_, t1_);
    return time;
  }
  float milliseconds() const { return time; }
  float seconds() const { return time * 1e-3f; }

 private:
  cudaEvent_t t0_, t1_;
};

}  // namespace util
}  // namespace sparsifyme

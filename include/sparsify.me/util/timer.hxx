// util/timer.hxx -- the stopwatch the operators and drivers bracket their phases with.
// Interface of the reference's util::timer_t (include/sparsify.me/util/timer.hxx:24-55): begin() and end()
// take the stream to record on and block the host until the event has happened, end() returns the interval
// in milliseconds, which also stays readable as `.time` / milliseconds() / seconds().
#pragma once
#include <cuda_runtime.h>

namespace sparsifyme {
namespace util {
namespace detail {

// one CUDA event that knows how to mark a point on a stream and wait for it
class marker {
 public:
  marker() { cudaEventCreate(&ev_); }
  ~marker() { cudaEventDestroy(ev_); }
  marker(const marker&) = delete;
  marker& operator=(const marker&) = delete;
  void drop(cudaStream_t s) const {
    cudaEventRecord(ev_, s);
    cudaEventSynchronize(ev_);
  }
  float ms_until(const marker& later) const {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_, later.ev_);
    return ms;
  }

 private:
  cudaEvent_t ev_ = nullptr;
};

}  // namespace detail

struct timer_t {
  float time = 0.f;  // last interval in milliseconds

  void begin(cudaStream_t stream = 0) { from_.drop(stream); }
  float end(cudaStream_t stream = 0) {
    to_.drop(stream);
    return time = from_.ms_until(to_);
  }
  float milliseconds() const { return time; }
  float seconds() const { return time / 1000.f; }

 private:
  detail::marker from_, to_;
};

}  // namespace util
}  // namespace sparsifyme

// api.cu -- library-level entry points of libsparsifyme_b200.so
#include "common.cuh"

extern "C" {

int spfy_version(void) { return 100; }  // 0.1.0

const char* spfy_last_error_string(void) { return spfy::err_buf(); }

uint64_t spfy_launch_count(void) { return spfy::launch_counter().load(std::memory_order_relaxed); }

}  // extern "C"

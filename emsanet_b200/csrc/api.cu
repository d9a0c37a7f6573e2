// Globals of the C ABI (error string, launch counter, version).
#include "../../include/emsanet_b200.h"
#include "common.h"

namespace eb {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace eb

extern "C" const char* eb200_last_error(void) { return eb::g_err; }
extern "C" int eb200_version(void) { return 100; }
extern "C" long long eb200_launch_count(void) { return eb::g_launches.load(); }

// zero `bytes` bytes on `stream` (scratch that must be cleared on the stream its consumer is launched on)
extern "C" int eb200_memset_zero(void* p, long long bytes, void* stream) {
  EB_REQUIRE(p && bytes >= 0, "eb200_memset_zero: bad argument");
  EB_CUDA(cudaMemsetAsync(p, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
  return 0;
}

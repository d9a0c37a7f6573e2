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

// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

namespace eb {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define EB_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) return ::eb::fail("%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define EB_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) return ::eb::fail(__VA_ARGS__); \
  } while (0)

inline int launch_check(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// Launch with programmatic stream serialization (PDL) and an optional cluster size.  EB200_NO_PDL=1 disables PDL.
// `kind` selects a bit of EB200_PDL_MASK (experiments): 1 halo conv, 2 halo wgrad, 4 generic conv/wgrad, 8 batch norm.
inline cudaError_t launch_ex(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args,
                             int cluster = 1, int kind = 1) {
  static int no_pdl = -1, mask = 7;   // measured (scripts/ab_pdl.sh): tensor-core kernels gain, the BN kernels lose
  if (no_pdl < 0) {
    no_pdl = getenv("EB200_NO_PDL") ? 1 : 0;
    if (const char* e = getenv("EB200_PDL_MASK")) mask = atoi(e);
  }
  const bool pdl = (mask & kind) != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && !no_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace eb

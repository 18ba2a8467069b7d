// Error plumbing and device queries shared by every entry point of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ure {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}
}  // namespace ure

extern "C" const char* ure_last_error(void) { return ure::g_err; }
extern "C" int ure_abi_version(void) { return URE_ABI_VERSION; }

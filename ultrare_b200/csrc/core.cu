// Error plumbing and device queries shared by every entry point of the C ABI.
#include <emmintrin.h>
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

// Host-side staging copy with non-temporal stores (the runtime's pinned staging pool, kernels.py).  The bytes go
// to DRAM without passing through the writing core's cache: when several threads fill different parts of a
// pinned buffer with ordinary stores, the GPU's DMA engine afterwards reads it at a third of the PCIe rate
// (dirty lines in many private caches; measured 18 vs 55 GB/s, tools/prof_upload.py).  dst 16-byte aligned.
extern "C" int ure_host_stage_copy(void* dst, const void* src, int64_t bytes) {
  using namespace ure;
  URE_REQUIRE((dst && src) || bytes == 0, URE_EINVAL, "ure_host_stage_copy: null argument");
  URE_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, URE_EINVAL, "ure_host_stage_copy: dst not 16-byte aligned");
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  int64_t i = 0;
  for (; i + 64 <= bytes; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
  }
  for (; i + 16 <= bytes; i += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)));
  if (i < bytes) memcpy(d + i, s + i, (size_t)(bytes - i));
  _mm_sfence();
  return 0;
}

// Asynchronous device -> page-locked host copy on the caller's stream (the host mirror's result read-back: one call
// per contiguous run instead of one framework dispatch per tensor).
extern "C" int ure_copy_to_host_async(void* h_dst, const void* d_src, int64_t bytes, void* stream) {
  using namespace ure;
  URE_REQUIRE(bytes >= 0 && (bytes == 0 || (h_dst && d_src)), URE_EINVAL, "ure_copy_to_host_async: bad argument");
  if (bytes == 0) return 0;
  URE_CUDA(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Asynchronous PAGE-LOCKED host -> device copy on the caller's stream (the uploads of RatingData: one driver call per
// array instead of a framework dispatch).
extern "C" int ure_copy_to_device_async(void* d_dst, const void* h_src, int64_t bytes, void* stream) {
  using namespace ure;
  URE_REQUIRE(bytes >= 0 && (bytes == 0 || (d_dst && h_src)), URE_EINVAL, "ure_copy_to_device_async: bad argument");
  if (bytes == 0) return 0;
  URE_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

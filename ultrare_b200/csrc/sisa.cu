// SISA bookkeeping on device: affected-shard routing and the owner-row merge.
//   routing  reference method/sisa.py:76-81   (nested Python loops, `user in list`)
//   merge    reference method/sisa.py:52-58 (learn), 107-113 (unlearn)
#include "common.cuh"

namespace ure {
namespace {

__global__ void route_kernel(const int32_t* __restrict__ owner, int n_user, const int32_t* __restrict__ del,
                             int n_del, int32_t* __restrict__ flags, int n_shards) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_del) return;
  const int u = del[j];
  if (u < 0 || u >= n_user) return;
  const int s = owner[u];
  if (s >= 0 && s < n_shards) flags[s] = 1;          // benign race: every writer stores 1
}

// one group of d/4 lanes per user row; pure row copy (bit exact)
__global__ void merge_kernel(const float* const* __restrict__ Pk, const int32_t* __restrict__ owner,
                             const int32_t* __restrict__ row_of, const int32_t* __restrict__ retrain,
                             float* __restrict__ merged, int n_user, int d4, int zero_unowned) {
  const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_user * d4;
  if (x >= total) return;
  const int u = (int)(x / d4);
  const int c = (int)(x % d4);
  const int s = owner[u];
  float4* dst = reinterpret_cast<float4*>(merged) + x;
  if (s < 0) {
    if (zero_unowned) *dst = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (retrain && retrain[s] == 0) return;
  const long long row = row_of ? row_of[u] : u;
  *dst = __ldg(reinterpret_cast<const float4*>(Pk[s]) + row * d4 + c);
}

__global__ void pack_f64_kernel(const double* __restrict__ cols, long long n, long long ld,
                                const int32_t* __restrict__ row_of, int n_map, int4* __restrict__ out) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    int u = (int)cols[j];
    const int it = (int)cols[ld + j];
    const float r = (float)cols[2 * ld + j];                // float32(float64): round to nearest even, as NumPy
    if (row_of) u = (u >= 0 && u < n_map) ? row_of[u] : -1;
    out[j] = make_int4(u, it, __float_as_int(r), 0);
  }
}

}  // namespace
}  // namespace ure

extern "C" int ure_pack_interactions_f64(const double* d_cols, int64_t n, int64_t ld, const int32_t* d_row_of,
                                         int32_t n_map, ure_inter_t* d_out, void* stream) {
  using namespace ure;
  URE_REQUIRE((d_cols && d_out) || n == 0, URE_EINVAL, "ure_pack_interactions_f64: null argument");
  URE_REQUIRE(ld >= n, URE_EINVAL, "ure_pack_interactions_f64: ld < n");
  if (n <= 0) return 0;
  const int blocks = (int)((n + 255) / 256 < (long long)(8 * num_sms()) ? (n + 255) / 256 : 8 * num_sms());
  pack_f64_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_cols, n, ld, d_row_of, n_map,
                                                                          reinterpret_cast<int4*>(d_out));
  URE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ure_route_deletions(const int32_t* d_owner, int32_t n_user, const int32_t* d_del, int32_t n_del,
                                   int32_t* d_flags, int32_t n_shards, void* stream) {
  using namespace ure;
  URE_REQUIRE(d_owner && d_flags && (d_del || n_del == 0), URE_EINVAL, "ure_route_deletions: null argument");
  if (n_del <= 0) return 0;
  route_kernel<<<(n_del + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_owner, n_user, d_del, n_del,
                                                                                 d_flags, n_shards);
  URE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ure_merge_user_rows(const float* const* d_P, const int32_t* d_owner, const int32_t* d_row_of,
                                   const int32_t* d_retrain, float* d_merged, int32_t n_user, int d,
                                   int zero_unowned, void* stream) {
  using namespace ure;
  URE_REQUIRE(d_P && d_owner && d_merged, URE_EINVAL, "ure_merge_user_rows: null argument");
  URE_REQUIRE(d > 0 && d % 4 == 0, URE_EUNSUPPORTED, "ure_merge_user_rows: d=%d must be a multiple of 4", d);
  if (n_user <= 0) return 0;
  const long long total = (long long)n_user * (d / 4);
  merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_P, d_owner, d_row_of, d_retrain, d_merged, n_user, d / 4, zero_unowned);
  URE_CUDA(cudaGetLastError());
  return 0;
}

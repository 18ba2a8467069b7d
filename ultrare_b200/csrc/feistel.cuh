// Keyed pseudo-random permutation of [0,n): 4-round Feistel network + cycle walking.
//
// Stands in for the per-epoch shuffle of DataLoader(shuffle=True) (reference
// read.py:133), which the reference draws from an unseeded torch generator and is
// therefore not reproducible.  Computed inline by the training
// kernel: no permutation array is ever stored or uploaded.  The CPU oracle restates
// the same function (oracle/mf.py: mix32 / perm_key / feistel_perm).
#pragma once
#include <stdint.h>

namespace ure {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__host__ __device__ __forceinline__ uint32_t perm_key(uint32_t seed, uint32_t shard, uint32_t epoch) {
  uint32_t k = mix32(seed ^ 0x9E3779B9u);
  k = mix32(k + shard * 0x85EBCA77u + 1u);
  k = mix32(k ^ (epoch * 0xC2B2AE3Du + 0x27D4EB2Fu));
  return k;
}

struct Feistel {
  uint32_t rk[4];
  uint32_t half, mask, n;

  __host__ __device__ void init(uint32_t n_, uint32_t key) {
    n = n_;
    uint32_t bits = 2;
    while (bits < 32 && (1u << bits) < n_) ++bits;     // bits = max(2, ceil(log2 n))
    half = (bits + 1) >> 1;
    mask = (1u << half) - 1u;
#pragma unroll
    for (int r = 0; r < 4; ++r) rk[r] = mix32(key + (uint32_t)r * 0x9E3779B9u);
  }

  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t j) const {
    if (n <= 1) return 0;
    uint32_t x = j;
    do {
      uint32_t L = x >> half, R = x & mask;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        uint32_t t = L ^ (mix32(R ^ rk[r]) & mask);
        L = R;
        R = t;
      }
      x = (L << half) | R;
    } while (x >= n);
    return x;
  }
};

}  // namespace ure

// Keyed pseudo-random permutation of [0,n): 6-round alternating (mixed-radix) Feistel network.
//
// Stands in for the per-epoch shuffle of DataLoader(shuffle=True) (reference read.py:133), which
// the reference draws from an unseeded torch generator and is therefore not reproducible
// (SURVEY.md §0.5, H8).  Computed inline by the training kernel: no permutation array is ever
// stored or uploaded.  The CPU oracle restates the same function (oracle/mf.py: mix32 / perm_key /
// feistel_perm).
//
// Domain: x = L*b + R with L in [0,a), R in [0,b), a = ceil(sqrt(n)), b = ceil(n/a), so a*b - n < a
// and cycle walking (re-encrypt while x >= n) almost never iterates -- no warp divergence.
//   even round r: L = (L + mulhi(mix32(R ^ rk[r]), a)) mod a
//   odd  round r: R = (R + mulhi(mix32(L ^ rk[r]), b)) mod b
// Each round is a bijection of [0,a) x [0,b); so is their composition, and so is its cycle-walked
// restriction to [0,n).
#pragma once
#include <math.h>
#include <stdint.h>

namespace ure {

constexpr int kFeistelRounds = 6;

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

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t x, uint32_t y) {
  return (uint32_t)(((uint64_t)x * (uint64_t)y) >> 32);
}

struct FeistelDomain {   // per shard, fixed for the whole training
  uint32_t n, a, b;
  __host__ __device__ void init(uint32_t n_) {
    n = n_;
    uint32_t r = (uint32_t)sqrt((double)n_);         // a = ceil(sqrt(n)): smallest a with a*a >= n
    while ((uint64_t)r * r < n_) ++r;
    while (r > 0 && (uint64_t)(r - 1) * (r - 1) >= n_) --r;
    a = r < 1 ? 1 : r;
    b = (n_ + a - 1) / a;
    if (b < 1) b = 1;
  }
};

struct FeistelKeys {     // per shard and epoch
  uint32_t rk[kFeistelRounds];
  __host__ __device__ void init(uint32_t key) {
#pragma unroll
    for (int r = 0; r < kFeistelRounds; ++r) rk[r] = mix32(key + (uint32_t)r * 0x9E3779B9u);
  }
};

__host__ __device__ __forceinline__ uint32_t feistel(const FeistelDomain& dm, const FeistelKeys& ks, uint32_t j) {
  if (dm.n <= 1) return 0;
  uint32_t x = j;
  do {
    uint32_t L = x / dm.b, R = x - L * dm.b;
#pragma unroll
    for (int r = 0; r < kFeistelRounds; ++r) {
      if ((r & 1) == 0) {
        L += mulhi32(mix32(R ^ ks.rk[r]), dm.a);
        if (L >= dm.a) L -= dm.a;
      } else {
        R += mulhi32(mix32(L ^ ks.rk[r]), dm.b);
        if (R >= dm.b) R -= dm.b;
      }
    }
    x = L * dm.b + R;
  } while (x >= dm.n);
  return x;
}

}  // namespace ure

// Keyed pseudo-random permutation of [0,n): 4-round alternating (mixed-radix) Feistel network.
//
// Stands in for the per-epoch shuffle of DataLoader(shuffle=True) (reference read.py:133), which
// the reference draws from an unseeded torch generator and is therefore not reproducible
// (SURVEY.md §0.5, H8).  Computed inline by the training kernels -- forward (position -> record) by the
// dense / lazy schedules, inverse (record -> position) by the owner schedule: no permutation array is
// ever stored or uploaded.  The CPU oracle restates the same function (oracle/mf.py: mix32 / perm_key /
// round_hash / feistel_perm).
//
// Domain: x = L*b + R with L in [0,a), R in [0,b), a = ceil(sqrt(n)), b = ceil(n/a), so a*b - n < a
// and cycle walking (re-encrypt while x >= n) almost never iterates -- no warp divergence.
//   even round r: L = (L + mulhi(round_hash(R ^ rk[r]), a)) mod a
//   odd  round r: R = (R + mulhi(round_hash(L ^ rk[r]), b)) mod b
// Each round is a bijection of [0,a) x [0,b); so is their composition, and so is its cycle-walked
// restriction to [0,n).  The round function is two multiplies and one xor-shift (the network is a batch
// shuffle, not a cipher); x / b uses a per-domain reciprocal.  About 45 instructions per evaluation.
#pragma once
#include <math.h>
#include <stdint.h>

namespace ure {

constexpr int kFeistelRounds = 4;

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {      // murmur3 finaliser: key schedule only
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__host__ __device__ __forceinline__ uint32_t round_hash(uint32_t x) {
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA6Bu;
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
  uint32_t n, a, b, rb;  // rb = floor(2^32 / b) (saturated): x / b = mulhi(x, rb) or one more
  __host__ __device__ void init(uint32_t n_) {
    n = n_;
    uint32_t r = (uint32_t)sqrt((double)n_);         // a = ceil(sqrt(n)): smallest a with a*a >= n
    while ((uint64_t)r * r < n_) ++r;
    while (r > 0 && (uint64_t)(r - 1) * (r - 1) >= n_) --r;
    a = r < 1 ? 1 : r;
    b = (n_ + a - 1) / a;
    if (b < 1) b = 1;
    const uint64_t q = 0x100000000ull / b;
    rb = q > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)q;
  }
  __host__ __device__ __forceinline__ void split(uint32_t x, uint32_t& L, uint32_t& R) const {
    L = mulhi32(x, rb);
    R = x - L * b;
    if (R >= b) { ++L; R -= b; }
  }
};

struct FeistelKeys {     // per shard and epoch
  uint32_t rk[kFeistelRounds];
  __host__ __device__ void init(uint32_t key) {
#pragma unroll
    for (int r = 0; r < kFeistelRounds; ++r) rk[r] = mix32(key + (uint32_t)r * 0x9E3779B9u);
  }
};

// position j of the epoch's visiting order -> record index
__host__ __device__ __forceinline__ uint32_t feistel(const FeistelDomain& dm, const FeistelKeys& ks, uint32_t j) {
  if (dm.n <= 1) return 0;
  uint32_t x = j;
  do {
    uint32_t L, R;
    dm.split(x, L, R);
#pragma unroll
    for (int r = 0; r < kFeistelRounds; ++r) {
      if ((r & 1) == 0) {
        L += mulhi32(round_hash(R ^ ks.rk[r]), dm.a);
        if (L >= dm.a) L -= dm.a;
      } else {
        R += mulhi32(round_hash(L ^ ks.rk[r]), dm.b);
        if (R >= dm.b) R -= dm.b;
      }
    }
    x = L * dm.b + R;
  } while (x >= dm.n);
  return x;
}

// record index -> its position in the epoch's visiting order; NI independent inversions interleaved
// (the rounds of one inversion are a single dependent chain)
template <int NI>
__host__ __device__ __forceinline__ void feistel_inverse_n(const FeistelDomain& dm, const FeistelKeys& ks,
                                                           uint32_t (&x)[NI], const bool (&live)[NI]) {
  if (dm.n <= 1) {
#pragma unroll
    for (int u = 0; u < NI; ++u) x[u] = 0;
    return;
  }
  bool pend[NI];
#pragma unroll
  for (int u = 0; u < NI; ++u) pend[u] = live[u];
  bool any = true;
  while (any) {
    uint32_t L[NI], R[NI];
#pragma unroll
    for (int u = 0; u < NI; ++u) dm.split(x[u], L[u], R[u]);
#pragma unroll
    for (int r = kFeistelRounds - 1; r >= 0; --r) {
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        if ((r & 1) == 0) {
          const uint32_t f = mulhi32(round_hash(R[u] ^ ks.rk[r]), dm.a);
          L[u] = L[u] >= f ? L[u] - f : L[u] + dm.a - f;
        } else {
          const uint32_t f = mulhi32(round_hash(L[u] ^ ks.rk[r]), dm.b);
          R[u] = R[u] >= f ? R[u] - f : R[u] + dm.b - f;
        }
      }
    }
    any = false;
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      if (pend[u]) {
        x[u] = L[u] * dm.b + R[u];
        pend[u] = x[u] >= dm.n;
        any |= pend[u];
      }
    }
  }
}

}  // namespace ure

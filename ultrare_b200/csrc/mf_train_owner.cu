// Owner-computes schedule of the fused MF training step (ure_mf_hparams_t::mode == URE_MF_OWNER).
//
// Same arithmetic as mf_train.cu (baseTrain, reference method/utils.py:58-98, + the dense
// optim.SGD(momentum, weight_decay) update of scratch.py:65-68), different data movement:
//
//   * Every CTA serves ONE shard and owns a contiguous slice of that shard's user rows and a slice of
//     its item rows.  The owned weights and momentum buffers live in shared memory for the whole launch.
//   * The shard's records exist twice, sorted by user and sorted by item (ure_mf_owner_prepare), so the
//     interactions of an owned row are one contiguous run of slots.  A batch is a random subset of the
//     shard (the per-epoch visiting order); which batch a slot belongs to in epoch e is
//     step_of(slot) = inverse_permutation_e(record index) / batch, evaluated once per epoch into a
//     shared-memory uint16 array -- for the NEXT epoch, slice by slice, in the shadow of the barrier.
//   * A step walks every owned row (in chunks of slots, handed out dynamically to groups of d/4 lanes):
//     the slots of the current batch gather the OTHER table's row (16-byte L2 loads), the error is
//     re-computed on both sides, the row gradient accumulates in registers, and the SGD update of the
//     row is applied right away (rows split over several chunks combine through shared memory).
//     No global atomics, no gradient arrays, no separate dense sweep.
//   * Updated rows are published to the buffer the NEXT step reads: P/Q and gP/gQ alternate
//     (reads of step j come from buffer j&1), so gradients always see pre-step weights (batch-synchronous
//     semantics of the reference) with ONE barrier per step -- and only among the CTAs of the shard.
#include "common.cuh"
#include "feistel.cuh"

namespace ure {
namespace {

constexpr int kOwnThreads = 1024;
constexpr int KMAX = URE_MAX_SHARDS;

struct OwnerWs {
  int need_smem;          // written by plan_kernel: dynamic shared memory the busiest CTA needs
  int avail_smem;         // what the launch can give
  int max_rows;           // max rows (user + item) per CTA
  int max_slots;          // max interactions (user side + item side) per CTA
  int planned_grid;       // grid size the plan was made for
  int planned_K;
  int pad[26];
  unsigned bar[KMAX][32]; // one barrier counter per shard, one 128-byte line each
};

struct Plan {
  int shard, q, c;        // my shard, my index among its c CTAs
  int ru0, ru1, ri0, ri1; // owned user rows / item rows
  int su0, mU, si0, mI;   // first slot and slot count in inter_u / inter_i
};

// first r in [0, n_rows] with off[r] >= target, moved down by one when that boundary is nearer
__device__ int nearest_boundary(const int32_t* off, int n_rows, long long target) {
  int lo = 0, hi = n_rows;                       // off[n_rows] = n >= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)off[mid] >= target) hi = mid; else lo = mid + 1;
  }
  if (lo > 0 && target - (long long)off[lo - 1] < (long long)off[lo] - target) --lo;
  return lo;
}

// CTA -> (shard, row slices).  CTAs are apportioned to shards by interaction count (largest remainder);
// inside a shard CTA q takes user rows up to the boundary nearest q*n/c, and item rows so that the
// CUMULATIVE (user + item) slot count of CTAs 0..q-1 is nearest 2*q*n/c (items are the finer grain).
// Block-cooperative (any block size >= 32, ends with a __syncthreads); scratch in shared memory.
struct PlanScratch {
  int c[KMAX];
  long long rem[KMAX];
  int ub[KMAX + 1], ib[KMAX + 1];
};

__device__ void make_plan(const ure_mf_shard_t* shards, int K, int cta, int n_cta, Plan& pl, PlanScratch& ps) {
  if (threadIdx.x == 0) {
    long long N = 0;
    for (int s = 0; s < K; ++s) N += shards[s].n;
    if (N < 1) N = 1;
    const long long spare = n_cta - K;
    int used = 0;
    for (int s = 0; s < K; ++s) {
      const long long x = spare * (long long)shards[s].n;
      ps.c[s] = 1 + (int)(x / N);
      ps.rem[s] = x % N;
      used += ps.c[s];
    }
    for (int left = n_cta - used; left > 0; --left) {
      int best = 0;
      for (int s = 1; s < K; ++s)
        if (ps.rem[s] > ps.rem[best]) best = s;
      ++ps.c[best];
      ps.rem[best] = -1;
    }
    int s = 0, base = 0;
    while (s < K - 1 && cta >= base + ps.c[s]) base += ps.c[s++];
    pl.shard = s; pl.q = cta - base; pl.c = ps.c[s];
  }
  __syncthreads();
  const ure_mf_shard_t& sh = shards[pl.shard];
  const int c = pl.c;
  const long long n = sh.n;
  // every boundary of the shard, in parallel; the item boundaries are made monotone by a running maximum
  for (int q = threadIdx.x; q <= c; q += blockDim.x) {
    int u = q <= 0 ? 0 : q >= c ? (int)sh.n_user : nearest_boundary(sh.off_u, sh.n_user, q * n / c);
    ps.ub[q] = u;
    ps.ib[q] = q <= 0 ? 0 : q >= c ? (int)sh.n_item
                                   : nearest_boundary(sh.off_i, sh.n_item, 2 * q * n / c - (long long)sh.off_u[u]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, i0 = 0, i1 = 0;
    for (int q = 0; q <= pl.q + 1; ++q) {
      run = max(run, ps.ib[q]);
      if (q == pl.q) i0 = run;
      if (q == pl.q + 1) i1 = run;
    }
    pl.ru0 = ps.ub[pl.q]; pl.ru1 = ps.ub[pl.q + 1];
    pl.ri0 = i0; pl.ri1 = i1;
    pl.su0 = sh.off_u[pl.ru0]; pl.mU = sh.off_u[pl.ru1] - pl.su0;
    pl.si0 = sh.off_i[pl.ri0]; pl.mI = sh.off_i[pl.ri1] - pl.si0;
  }
  __syncthreads();
}

__host__ __device__ inline int owner_smem_need(int rows, int slots, int d) {
  // w, buf, gacc [rows][d] fp32 | rowslot [rows+1], cpref [rows+1], done [rows] int | step_of [2][slots] u16
  long long b = (long long)rows * d * 12 + (long long)(rows + 1) * 8 + (long long)rows * 4 + 64;
  b += 2ll * 2 * ((slots + 7) & ~7);
  return b > 0x7fffffff ? 0x7fffffff : (int)b;
}

__global__ void plan_kernel(const ure_mf_shard_t* shards, int K, int d, OwnerWs* ws) {
  __shared__ PlanScratch ps;
  __shared__ Plan pl;
  make_plan(shards, K, blockIdx.x, gridDim.x, pl, ps);
  if (threadIdx.x == 0) {
    const int rows = (pl.ru1 - pl.ru0) + (pl.ri1 - pl.ri0);
    atomicMax(&ws->need_smem, owner_smem_need(rows, pl.mU + pl.mI, d));
    atomicMax(&ws->max_rows, rows);
    atomicMax(&ws->max_slots, pl.mU + pl.mI);
  }
}

// ---------------------------------------------------------------- set-up: counting sort of the records
__global__ void csr_count_kernel(const ure_mf_shard_t* shards) {
  const ure_mf_shard_t& sh = shards[blockIdx.y];
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < sh.n; j += (long long)gridDim.x * blockDim.x) {
    const int4 r = ld_stream_i4(sh.inter + j);
    atomicAdd(sh.off_u + r.x + 2, 1);
    atomicAdd(sh.off_i + r.y + 2, 1);
  }
}

// in-place inclusive scan of off[0 .. rows+2): one CTA per (shard, side)
__global__ void csr_scan_kernel(const ure_mf_shard_t* shards) {
  const ure_mf_shard_t& sh = shards[blockIdx.x >> 1];
  int32_t* off = (blockIdx.x & 1) ? sh.off_i : sh.off_u;
  const int len = ((blockIdx.x & 1) ? sh.n_item : sh.n_user) + 2;
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < len; base += blockDim.x) {
    const int x = base + threadIdx.x;
    int v = x < len ? off[x] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += a;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += a;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry + (warp > 0 ? s_warp[warp - 1] : 0);
    if (x < len) off[x] = v + carry;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_warp[31];
    __syncthreads();
  }
}

// after the scan off[r+1] = start of row r: used as the fill cursor, it ends as the start of row r+1
__global__ void csr_fill_kernel(const ure_mf_shard_t* shards) {
  const ure_mf_shard_t& sh = shards[blockIdx.y];
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < sh.n; j += (long long)gridDim.x * blockDim.x) {
    int4 r = ld_stream_i4(sh.inter + j);
    r.w = (int)j;
    const int pu = atomicAdd(sh.off_u + r.x + 1, 1);
    const int pi = atomicAdd(sh.off_i + r.y + 1, 1);
    reinterpret_cast<int4*>(sh.inter_u)[pu] = r;
    reinterpret_cast<int4*>(sh.inter_i)[pi] = r;
  }
}

__global__ void perm_inverse_kernel(const ure_mf_shard_t* shards, int epochs) {
  const ure_mf_shard_t& sh = shards[blockIdx.y];
  if (!sh.perm || !sh.perm_inv) return;
  const long long total = (long long)epochs * sh.n;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const long long e = x / sh.n;
    sh.perm_inv[e * sh.n + sh.perm[x]] = (int32_t)(x - e * sh.n);
  }
}

// ---------------------------------------------------------------- the inverse visiting order
__device__ __forceinline__ uint32_t feistel_inverse(const FeistelDomain& dm, const FeistelKeys& ks, uint32_t y) {
  if (dm.n <= 1) return 0;
  uint32_t x = y;
  do {
    uint32_t L = x / dm.b, R = x - L * dm.b;
#pragma unroll
    for (int r = kFeistelRounds - 1; r >= 0; --r) {
      if ((r & 1) == 0) {
        const uint32_t f = mulhi32(mix32(R ^ ks.rk[r]), dm.a);
        L = L >= f ? L - f : L + dm.a - f;
      } else {
        const uint32_t f = mulhi32(mix32(L ^ ks.rk[r]), dm.b);
        R = R >= f ? R - f : R + dm.b - f;
      }
    }
    x = L * dm.b + R;
  } while (x >= dm.n);
  return x;
}

// ---------------------------------------------------------------- the training kernel
template <int D>
__global__ void __launch_bounds__(kOwnThreads, 1)
mf_owner_kernel(const ure_mf_shard_t* __restrict__ shards, int K, ure_mf_hparams_t hp, int epochs,
                long long step_begin, long long step_end, OwnerWs* ws, unsigned dbg) {
  constexpr int G = D / 4;                 // lanes per row / interaction
  extern __shared__ __align__(16) unsigned char dyn[];
  constexpr int GPW = 32 / G;              // groups per warp
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ PlanScratch s_ps;
  __shared__ Plan s_pl;
  __shared__ ure_mf_shard_t s_sh;
  __shared__ int s_next;
  __shared__ float s_sse;

  const int tid = threadIdx.x, lane = tid & 31, gl = lane % G, gw = lane / G;
  make_plan(shards, K, blockIdx.x, gridDim.x, s_pl, s_ps);
  if (tid == 0) {
    s_sh = shards[s_pl.shard];
    s_next = 0;
    s_sse = 0.f;
  }
  __syncthreads();
  const Plan pl = s_pl;
  const ure_mf_shard_t& sh = s_sh;
  const int rowsU = pl.ru1 - pl.ru0, rowsI = pl.ri1 - pl.ri0, rows = rowsU + rowsI;
  const int m = pl.mU + pl.mI;
  const int B = hp.batch;
  const int n = sh.n;
  const int spe = (n + B - 1) / B;
  const long long t_end = spe > 0 ? min(step_end, (long long)spe * epochs) : step_begin;
  if (t_end <= step_begin) return;         // the whole shard (all of its CTAs) has nothing to do

  // ---- shared-memory carve-up
  float* const s_w = reinterpret_cast<float*>(dyn);
  float* const s_b = s_w + (size_t)rows * D;
  float* const s_g = s_b + (size_t)rows * D;
  int* const s_rowslot = reinterpret_cast<int*>(s_g + (size_t)rows * D);   // [rows+1] CTA-local first slot
  int* const s_cpref = s_rowslot + rows + 1;                                // [rows+1] chunk prefix
  int* const s_done = s_cpref + rows + 1;                                   // [rows]
  const int m_pad = (m + 7) & ~7;
  unsigned short* const s_step = reinterpret_cast<unsigned short*>(
      (reinterpret_cast<uintptr_t>(s_done + rows) + 15) & ~uintptr_t(15));  // [2][m_pad]

  // chunk length: a multiple of the 64-slot scan window, longer when batches are a small part of the epoch
  const int CL = 64 * max(1, min(8, (spe + 7) / 8));

  // ---- prologue: owned rows -> shared memory, slot offsets, chunk prefix
  for (int x = tid; x < rows * G; x += kOwnThreads) {
    const int r = x / G, c = x % G;
    const bool it = r >= rowsU;
    const size_t go = (size_t)(it ? pl.ri0 + (r - rowsU) : pl.ru0 + r) * D + 4 * c;
    const float4 w = ld_cg_f4((it ? sh.Q : sh.P) + go);
    const float4 b = ld_cg_f4((it ? sh.bufQ : sh.bufP) + go);
    *reinterpret_cast<float4*>(s_w + (size_t)r * D + 4 * c) = w;
    *reinterpret_cast<float4*>(s_b + (size_t)r * D + 4 * c) = b;
    *reinterpret_cast<float4*>(s_g + (size_t)r * D + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = tid; r <= rows; r += kOwnThreads) {
    s_rowslot[r] = r <= rowsU ? sh.off_u[pl.ru0 + r] - pl.su0
                              : pl.mU + sh.off_i[pl.ri0 + (r - rowsU)] - pl.si0;
    if (r < rows) s_done[r] = 0;
  }
  __syncthreads();
  if (tid < 32) {                          // chunk prefix: an empty row still gets one chunk (it must decay)
    int carry = 0;
    for (int base = 0; base < rows; base += 32) {
      const int r = base + lane;
      int v = 0;
      if (r < rows) v = max(1, (s_rowslot[r + 1] - s_rowslot[r] + CL - 1) / CL);
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
      }
      if (r < rows) s_cpref[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_cpref[rows] = carry;
  }
  __syncthreads();
  const int total_chunks = s_cpref[rows];

  // ---- visiting order -> step_of
  FeistelDomain dom;
  dom.init((uint32_t)n);
  const uint32_t magic = (uint32_t)(0x100000000ull / (uint32_t)B);     // floor(2^32/B): quotient low by <= 1
  auto fill_step_of = [&](int epoch, long long lo, long long hi) {
    unsigned short* out = s_step + (size_t)(epoch & 1) * m_pad;
    FeistelKeys ks;
    ks.init(perm_key(sh.perm_seed, (uint32_t)sh.shard_id, (uint32_t)epoch));
    const int32_t* pinv = sh.perm_inv ? sh.perm_inv + (long long)epoch * n : nullptr;
    for (long long sl = lo + tid; sl < hi; sl += kOwnThreads) {
      const ure_inter_t* rec = sl < pl.mU ? sh.inter_u + pl.su0 + sl : sh.inter_i + pl.si0 + (sl - pl.mU);
      const uint32_t j = (uint32_t)__ldg(&rec->pad);
      uint32_t pos;
      if (pinv) pos = (uint32_t)__ldg(pinv + j);
      else if (dbg & 2u) pos = j;
      else pos = feistel_inverse(dom, ks, j);
      uint32_t q = B == 1 ? pos : mulhi32(pos, magic);
      if ((q + 1) * (uint32_t)B <= pos) ++q;
      out[sl] = (unsigned short)q;
    }
  };
  int e = (int)(step_begin / spe), k = (int)(step_begin % spe);
  fill_step_of(e, 0, m);
  if (k > 0 && e + 1 < epochs) fill_step_of(e + 1, 0, (long long)k * m / spe);
  __syncthreads();

  auto lr_of = [&](int epoch) {
    double lr = (double)hp.lr0;
    for (int q = epoch / hp.lr_step; q > 0; --q) lr *= (double)hp.lr_decay;
    return (float)lr;
  };
  float nlr = -lr_of(e);
  const float wd = hp.weight_decay, mu = hp.momentum;
  unsigned* const counter = &ws->bar[pl.shard][0];
  unsigned bar_target = 0;
  double epoch_sse = 0.0;                  // thread 0 only

  for (long long t = step_begin; t < t_end; ++t) {
    const int rd = (int)((t - step_begin) & 1);
    const float* const Pr = rd ? sh.gP : sh.P;
    const float* const Qr = rd ? sh.gQ : sh.Q;
    float* const Pw = rd ? sh.P : sh.gP;
    float* const Qw = rd ? sh.Q : sh.gQ;
    const unsigned short* const stp = s_step + (size_t)(e & 1) * m_pad;
    float sse_l = 0.f;

    // ------------------------------------------------------------ owned rows, chunk by chunk
    // A warp takes GPW consecutive chunks from the CTA's queue, one per lane group; control flow is
    // warp-uniform (groups with fewer slots of this batch idle under predication).
    for (;;) {
      int x0 = 0;
      if (lane == 0) x0 = atomicAdd(&s_next, GPW);
      x0 = __shfl_sync(FULL, x0, 0);
      if (x0 >= total_chunks) break;
      const int x = x0 + gw;
      const bool have = x < total_chunks;
      int row = 0, nch = 1, sl0 = 0, sl1 = 0;
      if (have) {
        int lo = 0, hi = rows;             // largest row with cpref[row] <= x
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_cpref[mid] <= x) lo = mid; else hi = mid;
        }
        row = lo;
        nch = s_cpref[row + 1] - s_cpref[row];
        sl0 = s_rowslot[row] + (x - s_cpref[row]) * CL;
        sl1 = min(sl0 + CL, s_rowslot[row + 1]);
      }
      const bool it = row >= rowsU;
      const int4* const recs = reinterpret_cast<const int4*>(it ? sh.inter_i + pl.si0 - pl.mU : sh.inter_u + pl.su0);
      const float* const other = it ? Pr : Qr;
      const float4 wown = *reinterpret_cast<const float4*>(s_w + (size_t)row * D + 4 * gl);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

      for (int w0 = sl0;; w0 += 64) {
        const bool act = have && w0 < sl1;
        if (!__any_sync(FULL, act)) break;
        unsigned long long mask = 0;
        if (act) {
#pragma unroll
          for (int i = 0; i < 64 / G; ++i) {
            const int sl = w0 + gl + G * i;
            if (sl < sl1 && stp[sl] == (unsigned short)k) mask |= 1ull << (gl + G * i);
          }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mask |= __shfl_xor_sync(FULL, mask, o);
        while (__any_sync(FULL, mask != 0)) {
          int4 rec[4];
          float4 o4[4];
          bool ok[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ok[q] = mask != 0;
            rec[q] = make_int4(0, 0, 0, 0);
            if (ok[q]) {
              const int b = __ffsll((long long)mask) - 1;
              mask &= mask - 1;
              rec[q] = __ldg(recs + w0 + b);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[q]) o4[q] = ld_cg_f4(other + (size_t)(it ? rec[q].x : rec[q].y) * D + 4 * gl);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float dot = wown.x * o4[q].x;
            dot = fmaf(wown.y, o4[q].y, dot);
            dot = fmaf(wown.z, o4[q].z, dot);
            dot = fmaf(wown.w, o4[q].w, dot);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            const float err = ok[q] ? dot - __int_as_float(rec[q].z) : 0.f;
            const float ge = 2.f * err;
            acc.x = fmaf(ge, o4[q].x, acc.x); acc.y = fmaf(ge, o4[q].y, acc.y);
            acc.z = fmaf(ge, o4[q].z, acc.z); acc.w = fmaf(ge, o4[q].w, acc.w);
            if (!it && gl == 0) sse_l = fmaf(err, err, sse_l);
          }
        }
      }

      // ---------------------------------------------------------- the row's SGD update (whoever completes it)
      const bool multi = have && nch > 1;
      float* const gp = s_g + (size_t)row * D + 4 * gl;
      if (multi) {
        atomicAdd(gp + 0, acc.x); atomicAdd(gp + 1, acc.y); atomicAdd(gp + 2, acc.z); atomicAdd(gp + 3, acc.w);
      }
      __threadfence_block();
      __syncwarp();
      int old = 0;
      if (multi && gl == 0) old = atomicAdd(&s_done[row], 1);
      old = __shfl_sync(FULL, old, lane - gl);
      const bool finish = have && (!multi || old == nch - 1);
      if (multi && finish) {
        __threadfence_block();
        volatile float* vg = gp;
        acc = make_float4(vg[0], vg[1], vg[2], vg[3]);
        *reinterpret_cast<float4*>(gp) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gl == 0) s_done[row] = 0;
      }
      if (finish) {
        float* wp = s_w + (size_t)row * D + 4 * gl;
        float* bp = s_b + (size_t)row * D + 4 * gl;
        float4 w = *reinterpret_cast<float4*>(wp);
        float4 b = *reinterpret_cast<float4*>(bp);
        // torch SGD: d_p = g + wd*w (fma); buf = buf*mu + d_p; w = w + (-lr)*buf (fma)
        acc.x = fmaf(wd, w.x, acc.x); acc.y = fmaf(wd, w.y, acc.y);
        acc.z = fmaf(wd, w.z, acc.z); acc.w = fmaf(wd, w.w, acc.w);
        b.x = __fadd_rn(__fmul_rn(b.x, mu), acc.x); b.y = __fadd_rn(__fmul_rn(b.y, mu), acc.y);
        b.z = __fadd_rn(__fmul_rn(b.z, mu), acc.z); b.w = __fadd_rn(__fmul_rn(b.w, mu), acc.w);
        w.x = fmaf(nlr, b.x, w.x); w.y = fmaf(nlr, b.y, w.y); w.z = fmaf(nlr, b.z, w.z); w.w = fmaf(nlr, b.w, w.w);
        *reinterpret_cast<float4*>(wp) = w;
        *reinterpret_cast<float4*>(bp) = b;
        const size_t go = (size_t)(it ? pl.ri0 + (row - rowsU) : pl.ru0 + row) * D + 4 * gl;
        st_cg_f4((it ? Qw : Pw) + go, w);
      }
    }

    // ------------------------------------------------------------ loss, barrier among the shard's CTAs
    sse_l = warp_sum(sse_l);
    if (lane == 0 && sse_l != 0.f) atomicAdd(&s_sse, sse_l);
    __syncthreads();                       // every row of this CTA is updated and published
    if (tid == 0)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    bar_target += (unsigned)pl.c;
    const bool last_of_epoch = k + 1 == spe;
    // in the barrier's shadow: this step's slice of the NEXT epoch's step_of
    if (e + 1 < epochs) fill_step_of(e + 1, (long long)k * m / spe, (long long)(k + 1) * m / spe);
    if (tid == 0) {
      epoch_sse += (double)s_sse;
      s_sse = 0.f;
      if (last_of_epoch || t + 1 == t_end) {
        if (epoch_sse != 0.0) atomicAdd(sh.sse + e, epoch_sse);
        epoch_sse = 0.0;
      }
      s_next = 0;
      while (ld_acquire_u32(counter) < bar_target) {
      }
    }
    __syncthreads();
    if (last_of_epoch) { ++e; k = 0; nlr = -lr_of(e); }
    else ++k;
  }

  // ---- epilogue: the owned rows go back to P/Q (whatever the parity), momentum to bufP/bufQ, and the
  // alternate weight buffer gP/gQ is returned zeroed (the DENSE schedule's contract for its gradient scratch)
  for (int x = tid; x < rows * G; x += kOwnThreads) {
    const int r = x / G, c = x % G;
    const bool it = r >= rowsU;
    const size_t go = (size_t)(it ? pl.ri0 + (r - rowsU) : pl.ru0 + r) * D + 4 * c;
    st_cg_f4((it ? sh.Q : sh.P) + go, *reinterpret_cast<const float4*>(s_w + (size_t)r * D + 4 * c));
    st_cg_f4((it ? sh.bufQ : sh.bufP) + go, *reinterpret_cast<const float4*>(s_b + (size_t)r * D + 4 * c));
    st_cg_f4((it ? sh.gQ : sh.gP) + go, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

int max_dyn_smem(int* out) {
  int dev = 0, v = 0;
  URE_CUDA(cudaGetDevice(&dev));
  URE_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  *out = v - 8 * 1024;                     // the kernel's static shared memory (plan scratch, descriptor)
  return 0;
}

template <int D>
int launch_owner(const ure_mf_shard_t* d_shards, int K, const ure_mf_hparams_t& hp, int epochs, long long s0,
                 long long s1, OwnerWs* ws, int smem, unsigned dbg, cudaStream_t st) {
  auto kern = mf_owner_kernel<D>;
  URE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  URE_CUDA(cudaMemsetAsync(ws->bar, 0, sizeof(ws->bar), st));
  void* args[] = {(void*)&d_shards, (void*)&K,  (void*)&hp, (void*)&epochs,
                  (void*)&s0,       (void*)&s1, (void*)&ws, (void*)&dbg};
  URE_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(num_sms()), dim3(kOwnThreads), args, (size_t)smem, st));
  return 0;
}

unsigned g_owner_dbg = 0;

}  // namespace

int64_t mf_owner_workspace_bytes() { return (int64_t)sizeof(OwnerWs); }
void mf_owner_debug(unsigned flags) { g_owner_dbg = flags; }

// called by ure_mf_train when hparams.mode == URE_MF_OWNER.  h_need / h_avail: what ure_mf_owner_prepare
// planned, read back by the caller (the library never synchronises).
int mf_train_owner(const ure_mf_shard_t* d_shards, int n_shards, const ure_mf_hparams_t* hp, int epochs,
                   long long step_begin, long long step_end, void* d_workspace, cudaStream_t st) {
  URE_REQUIRE(n_shards <= num_sms(), URE_EUNSUPPORTED,
              "ure_mf_train(owner): %d shards need at least as many SMs (%d)", n_shards, num_sms());
  URE_REQUIRE(hp->owner_smem > 0, URE_EINVAL,
              "ure_mf_train(owner): hparams.owner_smem must carry the shared-memory bytes planned by "
              "ure_mf_owner_prepare (read back from the workspace)");
  int avail = 0;
  if (int rc = max_dyn_smem(&avail)) return rc;
  URE_REQUIRE(hp->owner_smem <= avail, URE_EUNSUPPORTED,
              "ure_mf_train(owner): the busiest CTA needs %d bytes of shared memory, %d available -- use the "
              "dense or lazy schedule for this problem size", hp->owner_smem, avail);
  if (step_end <= step_begin) return 0;
  auto* ws = static_cast<OwnerWs*>(d_workspace);
  const int smem = hp->owner_smem;
  switch (hp->d) {
    case 8: return launch_owner<8>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, smem, g_owner_dbg, st);
    case 16: return launch_owner<16>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, smem, g_owner_dbg, st);
    case 32: return launch_owner<32>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, smem, g_owner_dbg, st);
    case 64: return launch_owner<64>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, smem, g_owner_dbg, st);
    case 128: return launch_owner<128>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, smem, g_owner_dbg, st);
    default:
      set_error("ure_mf_train(owner): d=%d not in {8,16,32,64,128}", hp->d);
      return URE_EUNSUPPORTED;
  }
}

}  // namespace ure

extern "C" int ure_mf_owner_prepare(const ure_mf_shard_t* d_shards, int n_shards, const ure_mf_hparams_t* h_hp,
                                    int epochs, void* d_workspace, void* stream) {
  using namespace ure;
  URE_REQUIRE(d_shards && h_hp && d_workspace, URE_EINVAL, "ure_mf_owner_prepare: null argument");
  URE_REQUIRE(n_shards >= 1 && n_shards <= num_sms(), URE_EUNSUPPORTED,
              "ure_mf_owner_prepare: n_shards=%d outside [1,%d] (one CTA per shard at least)", n_shards, num_sms());
  auto st = static_cast<cudaStream_t>(stream);
  auto* ws = static_cast<OwnerWs*>(d_workspace);
  const int blocks = 2 * num_sms();
  csr_count_kernel<<<dim3(blocks, n_shards), 256, 0, st>>>(d_shards);
  csr_scan_kernel<<<2 * n_shards, 1024, 0, st>>>(d_shards);
  csr_fill_kernel<<<dim3(blocks, n_shards), 256, 0, st>>>(d_shards);
  perm_inverse_kernel<<<dim3(blocks, n_shards), 256, 0, st>>>(d_shards, epochs);
  int avail = 0;
  if (int rc = max_dyn_smem(&avail)) return rc;
  const int head[6] = {0, avail, 0, 0, num_sms(), n_shards};
  URE_CUDA(cudaMemcpyAsync(ws, head, sizeof(head), cudaMemcpyHostToDevice, st));
  plan_kernel<<<num_sms(), 32, 0, st>>>(d_shards, n_shards, h_hp->d, ws);
  URE_CUDA(cudaGetLastError());
  return 0;
}

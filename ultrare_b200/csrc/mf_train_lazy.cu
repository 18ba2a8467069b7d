// Lazy variant of the fused MF training step for tables far larger than a batch.
//
// The reference's optimiser is dense: optim.SGD(momentum, weight_decay) (scratch.py:65-68) updates EVERY
// row of both tables every step, so a row that receives no gradient still decays:
//     buf' = mu*buf + wd*w ;  w' = w - lr*buf'      i.e.   [w;buf] <- M [w;buf],
//     M = [[1 - lr*wd, -lr*mu], [wd, mu]].
// At 10 M users x 1 M items with 30 000-interaction batches a dense sweep costs ~40x the algorithmic
// traffic.  Here a row carries the number of steps already applied to it (`last`); when a step touches it,
// the skipped gradient-free steps are applied in closed form with M^n from a precomputed table (host,
// float64 -> float32), which is the same arithmetic in exact terms and within fp32 rounding in practice.
//
// Per step t (one cooperative launch for all steps, all shards; lr is constant over the launch):
//   gradients  gather (w, buf, last) of P[u] and Q[i], advance both rows to step t in registers,
//              dot / loss / red.global.add.v4.f32 into the (dense, otherwise zero) gradient rows, and
//              claim each touched row once (atomicOr on bit 31 of `last`) into the step's row list;
//   barrier
//   update     only the listed rows: advance to t, apply torch's SGD step with the gradient, store
//              w, buf, last = t+1, zero the gradient row;
//   barrier
// ure_mf_flush advances every row to the end of training before anything else reads the tables.
#include "common.cuh"
#include "feistel.cuh"

namespace ure {
namespace {

constexpr int kThreads = 1024;
constexpr int KM = URE_MAX_SHARDS;
constexpr unsigned kTouched = 0x80000000u;

struct LazyWorkspace {
  unsigned barrier;
  unsigned pad[63];
};

struct LazyPtrs {
  float* P; float* Q; float* bufP; float* bufQ; float* gP; float* gQ;
  int* lastP; int* lastQ; int* touched;
};

template <typename T>
__device__ __forceinline__ int find_segment(const T* prefix, int nseg, T x) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// row list of (step parity, table) inside ure_mf_shard_t::touched: [2][2][1 + batch]
__device__ __forceinline__ int* touched_list(int* base, int parity, int table, int batch) {
  return base + (size_t)(parity * 2 + table) * (size_t)(1 + batch);
}

__device__ __forceinline__ float4 advance_w(const float4 c, const float4 w, const float4 b) {
  return make_float4(fmaf(c.x, w.x, c.y * b.x), fmaf(c.x, w.y, c.y * b.y), fmaf(c.x, w.z, c.y * b.z),
                     fmaf(c.x, w.w, c.y * b.w));
}
__device__ __forceinline__ float4 advance_b(const float4 c, const float4 w, const float4 b) {
  return make_float4(fmaf(c.z, w.x, c.w * b.x), fmaf(c.z, w.y, c.w * b.y), fmaf(c.z, w.z, c.w * b.z),
                     fmaf(c.z, w.w, c.w * b.w));
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
mf_train_lazy_kernel(const ure_mf_shard_t* __restrict__ shards, int K, ure_mf_hparams_t hp, int epochs,
                     long long step_begin, long long step_end, LazyWorkspace* ws) {
  constexpr int G = D / 4;
  __shared__ LazyPtrs s_ptr[KM];
  __shared__ const ure_inter_t* s_inter[KM];
  __shared__ const int32_t* s_perm[KM];
  __shared__ double* s_sse[KM];
  __shared__ int s_n[KM], s_spe[KM], s_shard_id[KM];
  __shared__ uint32_t s_seed[KM];
  __shared__ FeistelDomain s_dom[KM];
  __shared__ FeistelKeys s_keys[KM];
  __shared__ int s_item_prefix[KM + 1], s_epoch[KM], s_start[KM];
  __shared__ int s_row_prefix[2 * KM + 1];
  __shared__ float s_sse_acc[KM];

  const int tid = threadIdx.x, lane = tid & 31, gl = lane % G;
  const float4* __restrict__ decay = reinterpret_cast<const float4*>(hp.decay);
  for (int s = tid; s < K; s += kThreads) {
    const ure_mf_shard_t sh = shards[s];
    s_ptr[s] = LazyPtrs{sh.P, sh.Q, sh.bufP, sh.bufQ, sh.gP, sh.gQ, sh.lastP, sh.lastQ, sh.touched};
    s_inter[s] = sh.inter; s_perm[s] = sh.perm; s_sse[s] = sh.sse;
    s_n[s] = sh.n;
    s_spe[s] = (sh.n + hp.batch - 1) / hp.batch;
    s_seed[s] = sh.perm_seed; s_shard_id[s] = sh.shard_id;
    s_dom[s].init((uint32_t)sh.n);
    s_sse_acc[s] = 0.f;
  }
  const long long n_threads = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + tid;
  const int n_warps = (int)(n_threads >> 5);
  // chunk index of this warp: CTA-minor, so that a step with fewer chunks than warps still uses every SM
  const int gwarp = (tid >> 5) * (int)gridDim.x + (int)blockIdx.x;
  const float wd = hp.weight_decay, mu = hp.momentum, nlr = -hp.lr0;
  unsigned bar_target = 0;
  __syncthreads();

  for (long long t = step_begin; t < step_end; ++t) {
    const int parity = (int)(t & 1);
    // ---- step tables (warp 0), lanes = shards
    if (tid < 32) {
      int carry = 0;
      for (int base = 0; base < K; base += 32) {
        const int s = base + lane;
        int cnt = 0, ep = -1, st = 0;
        if (s < K) {
          const int spe = s_spe[s];
          if (spe > 0 && t < (long long)spe * epochs) {
            ep = (int)(t / spe);
            st = (int)(t % spe) * hp.batch;
            cnt = min(hp.batch, s_n[s] - st);
            if (st == 0 || t == step_begin) s_keys[s].init(perm_key(s_seed[s], (uint32_t)s_shard_id[s], (uint32_t)ep));
          }
          s_epoch[s] = ep;
          s_start[s] = st;
        }
        int v = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += a; }
        if (s < K) s_item_prefix[s + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      if (lane == 0) s_item_prefix[0] = 0;
    }
    __syncthreads();

    // ---------------------------------------------------------------- gradients
    const int total = s_item_prefix[K];
    const int tt = (int)t;                      // shard-local step count == global step (all shards start at 0)
    for (int wc = gwarp; wc * 32 < total; wc += n_warps) {
      const int item = wc * 32 + lane;
      int s = -1, u = 0, it = 0;
      float r = 0.f;
      if (item < total) {
        s = find_segment(s_item_prefix, K, item);
        const int j = s_start[s] + (item - s_item_prefix[s]);
        const int32_t* pm = s_perm[s];
        const uint32_t idx = pm ? (uint32_t)__ldg(pm + (long long)s_epoch[s] * s_n[s] + j)
                                : feistel(s_dom[s], s_keys[s], (uint32_t)j);
        const int4 rec = ld_stream_i4(s_inter[s] + idx);
        u = rec.x; it = rec.y; r = __int_as_float(rec.z);
      }
      float my_e = 0.f;
#pragma unroll 1
      for (int q = 0; q < G; ++q) {             // the group's G interactions, one at a time (4 row gathers each)
        const int uq = __shfl_sync(0xffffffffu, u, q, G);
        const int iq = __shfl_sync(0xffffffffu, it, q, G);
        const float rq = __shfl_sync(0xffffffffu, r, q, G);
        const int sq = __shfl_sync(0xffffffffu, s, q, G);
        float dot = 0.f;
        float4 pu = make_float4(0.f, 0.f, 0.f, 0.f), qi = pu;
        LazyPtrs tp;
        if (sq >= 0) {
          tp = s_ptr[sq];
          const size_t ou = (size_t)uq * D + 4 * gl, oi = (size_t)iq * D + 4 * gl;
          const float4 wu = ld_cg_f4(tp.P + ou), bu = ld_cg_f4(tp.bufP + ou);
          const float4 wi = ld_cg_f4(tp.Q + oi), bi = ld_cg_f4(tp.bufQ + oi);
          const int lu = (int)(__ldcg(tp.lastP + uq) & 0x7fffffff), li = (int)(__ldcg(tp.lastQ + iq) & 0x7fffffff);
          pu = advance_w(__ldg(decay + (tt - lu)), wu, bu);
          qi = advance_w(__ldg(decay + (tt - li)), wi, bi);
          dot = pu.x * qi.x;
          dot = fmaf(pu.y, qi.y, dot);
          dot = fmaf(pu.z, qi.z, dot);
          dot = fmaf(pu.w, qi.w, dot);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o, G);
        const float e = dot - rq;
        if (gl == q) my_e = e;
        if (sq >= 0) {
          const float ge = 2.f * e;
          red_add_f4(tp.gP + (size_t)uq * D + 4 * gl, ge * qi.x, ge * qi.y, ge * qi.z, ge * qi.w);
          red_add_f4(tp.gQ + (size_t)iq * D + 4 * gl, ge * pu.x, ge * pu.y, ge * pu.z, ge * pu.w);
          if (gl == 0) {                         // claim each touched row once for the update phase
            if (!((unsigned)atomicOr(tp.lastP + uq, (int)kTouched) & kTouched)) {
              int* lst = touched_list(tp.touched, parity, 0, hp.batch);
              lst[1 + atomicAdd(lst, 1)] = uq;
            }
            if (!((unsigned)atomicOr(tp.lastQ + iq, (int)kTouched) & kTouched)) {
              int* lst = touched_list(tp.touched, parity, 1, hp.batch);
              lst[1 + atomicAdd(lst, 1)] = iq;
            }
          }
        }
      }
      const float e2 = s >= 0 ? my_e * my_e : 0.f;
      const int s0 = __shfl_sync(0xffffffffu, s, 0);
      if (__all_sync(0xffffffffu, s == s0 || s < 0)) {
        const float wsum = warp_sum(e2);
        if (lane == 0 && s0 >= 0) atomicAdd(&s_sse_acc[s0], wsum);
      } else if (s >= 0) {
        atomicAdd(&s_sse_acc[s], e2);
      }
    }
    __syncthreads();
    for (int s = tid; s < K; s += kThreads) {
      const float v = s_sse_acc[s];
      if (v != 0.f) { atomicAdd(s_sse[s] + s_epoch[s], (double)v); s_sse_acc[s] = 0.f; }
    }
    grid_barrier(&ws->barrier, bar_target);

    // ---------------------------------------------------------------- update of the touched rows
    if (tid < 32) {                              // flattened (shard, table, list entry) space
      int carry = 0;
      for (int base = 0; base < 2 * K; base += 32) {
        const int seg = base + lane;
        int cnt = 0;
        if (seg < 2 * K && s_epoch[seg >> 1] >= 0)
          cnt = __ldcg(touched_list(s_ptr[seg >> 1].touched, parity, seg & 1, hp.batch));
        int v = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += a; }
        if (seg < 2 * K) s_row_prefix[seg + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      if (lane == 0) s_row_prefix[0] = 0;
    }
    __syncthreads();
    const long long rows = s_row_prefix[2 * K];
    for (long long xb = gtid - lane; xb < rows * G; xb += n_threads) {      // warp-uniform trip count
      const long long x = xb + lane;
      const bool live = x < rows * G;
      const int e = (int)((live ? x : xb) / G), c = (int)(x % G);
      const int seg = find_segment(s_row_prefix, 2 * K, e);
      const LazyPtrs tp = s_ptr[seg >> 1];
      const int* lst = touched_list(tp.touched, parity, seg & 1, hp.batch);
      const int row = __ldcg(lst + 1 + (e - s_row_prefix[seg]));
      float* W = (seg & 1) ? tp.Q : tp.P;
      float* Bf = (seg & 1) ? tp.bufQ : tp.bufP;
      float* Gr = (seg & 1) ? tp.gQ : tp.gP;
      int* last = (seg & 1) ? tp.lastQ : tp.lastP;
      const size_t off = (size_t)row * D + 4 * c;
      const int l = (int)(__ldcg(last + row) & 0x7fffffff);
      const float4 cf = __ldg(decay + (tt - l));
      const float4 w0 = ld_cg_f4(W + off), b0 = ld_cg_f4(Bf + off);
      float4 w = advance_w(cf, w0, b0), b = advance_b(cf, w0, b0), g = ld_cg_f4(Gr + off);
      g.x = fmaf(wd, w.x, g.x); g.y = fmaf(wd, w.y, g.y); g.z = fmaf(wd, w.z, g.z); g.w = fmaf(wd, w.w, g.w);
      b.x = __fadd_rn(__fmul_rn(b.x, mu), g.x); b.y = __fadd_rn(__fmul_rn(b.y, mu), g.y);
      b.z = __fadd_rn(__fmul_rn(b.z, mu), g.z); b.w = __fadd_rn(__fmul_rn(b.w, mu), g.w);
      w.x = fmaf(nlr, b.x, w.x); w.y = fmaf(nlr, b.y, w.y); w.z = fmaf(nlr, b.z, w.z); w.w = fmaf(nlr, b.w, w.w);
      if (live) {
        st_cg_f4(W + off, w);
        st_cg_f4(Bf + off, b);
        st_cg_f4(Gr + off, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      __syncwarp();                              // all G lanes of the row have read `last` before it is rewritten
      if (live && c == 0) __stcg(last + row, tt + 1);    // also clears the touched flag
    }
    // the other parity's counters were consumed in step t-1: zero them for step t+1
    if (blockIdx.x == 0)
      for (int seg = tid; seg < 2 * K; seg += kThreads)
        *touched_list(s_ptr[seg >> 1].touched, parity ^ 1, seg & 1, hp.batch) = 0;
    grid_barrier(&ws->barrier, bar_target);
  }
}

// every row with fewer than `target` steps applied is advanced in closed form
__global__ void flush_kernel(float* __restrict__ W, float* __restrict__ Bf, int* __restrict__ last, long long n_rows,
                             int d4, int target, const float4* __restrict__ decay) {
  const long long total = n_rows * d4;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const long long row = x / d4;
    const int l = last[row] & 0x7fffffff;
    if (l >= target) continue;
    const float4 cf = __ldg(decay + (target - l));
    float4* wp = reinterpret_cast<float4*>(W) + x;
    float4* bp = reinterpret_cast<float4*>(Bf) + x;
    const float4 w = *wp, b = *bp;
    *wp = advance_w(cf, w, b);
    *bp = advance_b(cf, w, b);
  }
}
__global__ void set_last_kernel(int* __restrict__ last, long long n_rows, int target) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x)
    if ((last[r] & 0x7fffffff) < target) last[r] = target;
}

template <int D>
int launch_lazy(const ure_mf_shard_t* d_shards, int K, const ure_mf_hparams_t& hp, int epochs, long long s0,
                long long s1, LazyWorkspace* ws, cudaStream_t st) {
  auto kern = mf_train_lazy_kernel<D>;
  int occ = 0;
  URE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0));
  URE_REQUIRE(occ >= 1, URE_ECOOP, "mf_train_lazy_kernel<%d> cannot be resident", D);
  const int grid = num_sms();
  URE_CUDA(cudaMemsetAsync(&ws->barrier, 0, sizeof(unsigned), st));
  void* args[] = {(void*)&d_shards, (void*)&K, (void*)&hp, (void*)&epochs, (void*)&s0, (void*)&s1, (void*)&ws};
  URE_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kThreads), args, 0, st));
  return 0;
}

}  // namespace

// called by ure_mf_train when hparams.lazy != 0
int mf_train_lazy(const ure_mf_shard_t* d_shards, int n_shards, const ure_mf_hparams_t* hp, int epochs,
                  long long step_begin, long long step_end, void* d_workspace, cudaStream_t st) {
  URE_REQUIRE(hp->decay && hp->decay_len > step_end, URE_EINVAL,
              "ure_mf_train(lazy): decay table must cover %lld steps (has %d)", (long long)step_end, hp->decay_len);
  URE_REQUIRE(epochs <= hp->lr_step, URE_EUNSUPPORTED,
              "ure_mf_train(lazy): the learning rate must be constant over the training (epochs <= lr_step)");
  auto* ws = static_cast<LazyWorkspace*>(d_workspace);
  switch (hp->d) {
    case 8: return launch_lazy<8>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, st);
    case 16: return launch_lazy<16>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, st);
    case 32: return launch_lazy<32>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, st);
    case 64: return launch_lazy<64>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, st);
    case 128: return launch_lazy<128>(d_shards, n_shards, *hp, epochs, step_begin, step_end, ws, st);
    default:
      set_error("ure_mf_train: d=%d not in {8,16,32,64,128}", hp->d);
      return URE_EUNSUPPORTED;
  }
}

int mf_flush_lazy(const ure_mf_shard_t* h_shards, int n_shards, const ure_mf_hparams_t* hp, int epochs,
                  long long step_now, cudaStream_t st) {
  URE_REQUIRE(hp->decay && h_shards, URE_EINVAL, "ure_mf_flush: null argument");
  const float4* decay = reinterpret_cast<const float4*>(hp->decay);
  const int d4 = hp->d / 4;
  const int cap = num_sms() * 8;
  for (int s = 0; s < n_shards; ++s) {
    const ure_mf_shard_t& sh = h_shards[s];
    const long long spe = (sh.n + hp->batch - 1) / hp->batch;
    const long long done = spe * epochs < step_now ? spe * epochs : step_now;
    URE_REQUIRE(done < hp->decay_len, URE_EINVAL, "ure_mf_flush: decay table too short");
    struct { float* W; float* B; int* last; long long rows; } tabs[2] = {
        {sh.P, sh.bufP, sh.lastP, sh.n_user}, {sh.Q, sh.bufQ, sh.lastQ, sh.n_item}};
    for (auto& tb : tabs) {
      if (tb.rows <= 0) continue;
      long long blocks = (tb.rows * d4 + 255) / 256;
      if (blocks > cap) blocks = cap;
      flush_kernel<<<(unsigned)blocks, 256, 0, st>>>(tb.W, tb.B, tb.last, tb.rows, d4, (int)done, decay);
      long long b2 = (tb.rows + 255) / 256;
      if (b2 > cap) b2 = cap;
      set_last_kernel<<<(unsigned)b2, 256, 0, st>>>(tb.last, tb.rows, (int)done);
    }
  }
  URE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ure

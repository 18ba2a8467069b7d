// Fused Matrix-Factorization training step for ALL shard models in one persistent launch.
//
// Replaces the body of baseTrain (reference method/utils.py:58-98): per batch
//   pred = sum_t P[u,t]*Q[i,t]; e = pred - r; L += e^2;
//   gP[u] += 2e*Q[i]; gQ[i] += 2e*P[u]            (all with PRE-step weights)
// followed by the dense optim.SGD(momentum, weight_decay) update of scratch.py:65-68
// on every row of both tables, for every shard (the K sequential Scratch.train calls
// of method/sisa.py:33-36,86-89 become one launch).
//
// Structure per global step t (cooperative launch, one CTA per SM):
//   phase A  all shards' batches, flattened and split evenly over the grid; a group of
//            d/4 lanes owns one interaction: 16-byte gathers of P[u], Q[i], shuffle
//            dot product, red.global.add.v4.f32 scatter of both gradient rows;
//   barrier
//   phase B  dense sweep over all rows of all active shards: g += wd*w; buf = mu*buf+g;
//            w -= lr*buf; g = 0   (in place);
//   barrier
// The visiting order is either an explicit permutation (parity runs against the
// reference) or the inline Feistel permutation (feistel.cuh).
#include "common.cuh"
#include "feistel.cuh"

namespace ure {
namespace {

constexpr int kThreads = 1024;

struct Workspace {
  unsigned barrier;
  unsigned pad[63];
};

// Per-CTA shard tables (static shared memory, URE_MAX_SHARDS entries): direct LDS addressing.
struct ShardPtrs {      // 32 bytes: fetched with two LDS.128
  float* P; float* Q; float* gP; float* gQ;
};

// largest s with prefix[s] <= x  (prefix[0] = 0, prefix[K] = total > x)
template <typename T>
__device__ __forceinline__ int find_segment(const T* prefix, int nseg, T x) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// Gradient work of one 32-interaction warp chunk.  Lane l fetched interaction l (u, it, r, shard s);
// a group of G = D/4 lanes processes its G interactions SUB at a time (2*SUB 16-byte gathers in flight).
// UNIFORM: the whole chunk lies in shard tables `tp`; otherwise each interaction looks its shard up.
template <int D, bool UNIFORM>
__device__ __forceinline__ float chunk_gradients(int u, int it, float r, int s, const ShardPtrs& tp,
                                                  const ShardPtrs* s_ptr, int gl) {
  constexpr int G = D / 4;
  constexpr int SUB = (G < 4) ? G : 4;
  float my_e = 0.f;
#pragma unroll
  for (int q0 = 0; q0 < G; q0 += SUB) {
    float4 pu[SUB], qi[SUB];
    int uq[SUB], iq[SUB], sq[SUB];
    float rq[SUB];
#pragma unroll
    for (int q = 0; q < SUB; ++q) {
      uq[q] = __shfl_sync(0xffffffffu, u, q0 + q, G);
      iq[q] = __shfl_sync(0xffffffffu, it, q0 + q, G);
      rq[q] = __shfl_sync(0xffffffffu, r, q0 + q, G);
      sq[q] = __shfl_sync(0xffffffffu, s, q0 + q, G);
      pu[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      qi[q] = pu[q];
      if (sq[q] >= 0) {
        const float* Pb = UNIFORM ? tp.P : s_ptr[sq[q]].P;
        const float* Qb = UNIFORM ? tp.Q : s_ptr[sq[q]].Q;
        pu[q] = ld_cg_f4(Pb + (size_t)uq[q] * D + 4 * gl);
        qi[q] = ld_cg_f4(Qb + (size_t)iq[q] * D + 4 * gl);
      }
    }
#pragma unroll
    for (int q = 0; q < SUB; ++q) {
      float dot = pu[q].x * qi[q].x;
      dot = fmaf(pu[q].y, qi[q].y, dot);
      dot = fmaf(pu[q].z, qi[q].z, dot);
      dot = fmaf(pu[q].w, qi[q].w, dot);
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o, G);
      const float e = dot - rq[q];
      if (gl == q0 + q) my_e = e;
      if (sq[q] >= 0) {
        float* gPb = UNIFORM ? tp.gP : s_ptr[sq[q]].gP;
        float* gQb = UNIFORM ? tp.gQ : s_ptr[sq[q]].gQ;
        const float ge = 2.f * e;
        red_add_f4(gPb + (size_t)uq[q] * D + 4 * gl, ge * qi[q].x, ge * qi[q].y, ge * qi[q].z, ge * qi[q].w);
        red_add_f4(gQb + (size_t)iq[q] * D + 4 * gl, ge * pu[q].x, ge * pu[q].y, ge * pu[q].z, ge * pu[q].w);
      }
    }
  }
  return my_e;
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
mf_train_kernel(const ure_mf_shard_t* __restrict__ shards, int K, ure_mf_hparams_t hp, int epochs,
                long long step_begin, long long step_end, Workspace* ws) {
  constexpr int G = D / 4;                 // lanes per interaction
  constexpr int KM = URE_MAX_SHARDS;
  __shared__ ShardPtrs s_ptr[KM];
  __shared__ const ure_inter_t* s_inter[KM];
  __shared__ const int32_t* s_perm[KM];
  __shared__ float* s_buf[2 * KM];         // bufP, bufQ
  __shared__ double* s_sse[KM];
  __shared__ long long s_row_prefix[2 * KM + 1];
  __shared__ int s_n[KM], s_nuser[KM], s_nitem[KM], s_spe[KM], s_shard_id[KM];
  __shared__ uint32_t s_seed[KM];
  __shared__ int s_item_prefix[KM + 1], s_epoch[KM], s_start[KM];
  __shared__ float s_lr[KM], s_sse_acc[KM];
  __shared__ Feistel s_fe[KM];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int gl = lane % G;
  for (int s = tid; s < K; s += kThreads) {
    const ure_mf_shard_t sh = shards[s];
    s_ptr[s] = ShardPtrs{sh.P, sh.Q, sh.gP, sh.gQ};
    s_inter[s] = sh.inter; s_perm[s] = sh.perm;
    s_buf[2 * s] = sh.bufP; s_buf[2 * s + 1] = sh.bufQ;
    s_sse[s] = sh.sse;
    s_n[s] = sh.n; s_nuser[s] = sh.n_user; s_nitem[s] = sh.n_item;
    s_spe[s] = (sh.n + hp.batch - 1) / hp.batch;
    s_seed[s] = sh.perm_seed; s_shard_id[s] = sh.shard_id;
    s_sse_acc[s] = 0.f;
  }
  unsigned bar_target = 0;
  const long long n_threads = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + tid;
  const int n_warps = (int)(n_threads >> 5);
  const int gwarp = (int)(gtid >> 5);
  __syncthreads();

  for (long long t = step_begin; t < step_end; ++t) {
    // ---------------------------------------------------------------- per-step tables
    for (int s = tid; s < K; s += kThreads) {
      const int spe = s_spe[s];
      const bool active = spe > 0 && t < (long long)spe * epochs;
      int ep = 0, cnt = 0, st = 0;
      if (active) {
        ep = (int)(t / spe);
        st = (int)(t % spe) * hp.batch;
        cnt = min(hp.batch, s_n[s] - st);
        s_fe[s].init((uint32_t)s_n[s], perm_key(s_seed[s], (uint32_t)s_shard_id[s], (uint32_t)ep));
        double lr = (double)hp.lr0;
        for (int q = ep / hp.lr_step; q > 0; --q) lr *= (double)hp.lr_decay;
        s_lr[s] = (float)lr;
      }
      s_epoch[s] = active ? ep : -1;
      s_start[s] = st;
      s_item_prefix[s + 1] = cnt;                               // counts, scanned below
      s_row_prefix[2 * s + 1] = active ? (long long)s_nuser[s] * G : 0;
      s_row_prefix[2 * s + 2] = active ? (long long)s_nitem[s] * G : 0;
    }
    __syncthreads();
    if (tid < 32) {                                              // warp 0: inclusive scans
      int carry = 0;
      for (int base = 0; base < K; base += 32) {
        int idx = base + lane;
        int v = idx < K ? s_item_prefix[idx + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (idx < K) s_item_prefix[idx + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      long long carry2 = 0;
      for (int base = 0; base < 2 * K; base += 32) {
        int idx = base + lane;
        long long v = idx < 2 * K ? s_row_prefix[idx + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (idx < 2 * K) s_row_prefix[idx + 1] = v + carry2;
        carry2 += __shfl_sync(0xffffffffu, v, 31);
      }
      if (lane == 0) { s_item_prefix[0] = 0; s_row_prefix[0] = 0; }
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase A: gradients
    const int total = s_item_prefix[K];
    float acc = 0.f;          // sum of e^2 of this lane's own interactions, all in shard acc_s
    int acc_s = -1;
    for (int wc = gwarp; wc * 32 < total; wc += n_warps) {
      const int item = wc * 32 + lane;
      const bool valid = item < total;
      int s = -1;
      int u = 0, it = 0;
      float r = 0.f;
      if (valid) {
        s = find_segment(s_item_prefix, K, item);
        const int j = s_start[s] + (item - s_item_prefix[s]);
        const int32_t* pm = s_perm[s];
        const uint32_t idx = pm ? (uint32_t)__ldg(pm + (long long)s_epoch[s] * s_n[s] + j)
                                : s_fe[s]((uint32_t)j);
        const int4 rec = ld_stream_i4(s_inter[s] + idx);
        u = rec.x; it = rec.y; r = __int_as_float(rec.z);
      }
      // the common case: the whole 32-interaction chunk lies in one shard -> table pointers once per warp
      const int s0 = __shfl_sync(0xffffffffu, s, 0);
      const bool uniform = __all_sync(0xffffffffu, s == s0 || !valid);
      const ShardPtrs tp = s_ptr[s0];
      const float my_e = uniform ? chunk_gradients<D, true>(u, it, r, s, tp, s_ptr, gl)
                                 : chunk_gradients<D, false>(u, it, r, s, tp, s_ptr, gl);
      // loss: warp-reduce when the chunk is single-shard (one shared atomic per warp, not per lane)
      const float e2 = valid ? my_e * my_e : 0.f;
      if (uniform) {
        const float w = warp_sum(e2);
        if (s0 != acc_s) {
          if (acc_s >= 0 && lane == 0) atomicAdd(&s_sse_acc[acc_s], acc);
          acc = 0.f;
          acc_s = s0;
        }
        acc += w;                                   // every lane carries the warp total; lane 0 publishes
      } else if (valid) {
        atomicAdd(&s_sse_acc[s], e2);
      }
    }
    if (acc_s >= 0 && lane == 0) atomicAdd(&s_sse_acc[acc_s], acc);
    __syncthreads();
    for (int s = tid; s < K; s += kThreads) {
      const float v = s_sse_acc[s];
      if (v != 0.f) {
        atomicAdd(s_sse[s] + s_epoch[s], (double)v);
        s_sse_acc[s] = 0.f;
      }
    }
    grid_barrier(&ws->barrier, bar_target);

    // ---------------------------------------------------------------- phase B: dense SGD sweep
    const long long total4 = s_row_prefix[2 * K];
    const float wd = hp.weight_decay, mu = hp.momentum;
    for (long long x = gtid; x < total4; x += n_threads) {
      const int seg = find_segment(s_row_prefix, 2 * K, x);
      const size_t off = (size_t)(x - s_row_prefix[seg]) * 4;
      const int s = seg >> 1;
      const ShardPtrs tp = s_ptr[s];
      float* W = (seg & 1) ? tp.Q : tp.P;
      float* Gr = (seg & 1) ? tp.gQ : tp.gP;
      float* Bf = s_buf[seg];
      const float nlr = -s_lr[s];
      float4 g = ld_cg_f4(Gr + off), w = ld_cg_f4(W + off), b = ld_cg_f4(Bf + off);
      // torch SGD: d_p = g + wd*w (fma); buf = buf*mu + d_p; w = w + (-lr)*buf (fma)
      g.x = fmaf(wd, w.x, g.x); g.y = fmaf(wd, w.y, g.y); g.z = fmaf(wd, w.z, g.z); g.w = fmaf(wd, w.w, g.w);
      b.x = __fadd_rn(__fmul_rn(b.x, mu), g.x); b.y = __fadd_rn(__fmul_rn(b.y, mu), g.y);
      b.z = __fadd_rn(__fmul_rn(b.z, mu), g.z); b.w = __fadd_rn(__fmul_rn(b.w, mu), g.w);
      w.x = fmaf(nlr, b.x, w.x); w.y = fmaf(nlr, b.y, w.y); w.z = fmaf(nlr, b.z, w.z); w.w = fmaf(nlr, b.w, w.w);
      st_cg_f4(W + off, w);
      st_cg_f4(Bf + off, b);
      st_cg_f4(Gr + off, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    grid_barrier(&ws->barrier, bar_target);
  }
}

template <int D>
int launch(const ure_mf_shard_t* d_shards, int K, const ure_mf_hparams_t& hp, int epochs, long long s0,
           long long s1, Workspace* ws, cudaStream_t st) {
  auto kern = mf_train_kernel<D>;
  const size_t smem = 0;
  int occ = 0;
  URE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
  URE_REQUIRE(occ >= 1, URE_ECOOP, "mf_train_kernel<%d> cannot be resident (smem %zu)", D, smem);
  const int grid = num_sms();
  URE_CUDA(cudaMemsetAsync(&ws->barrier, 0, sizeof(unsigned), st));
  void* args[] = {(void*)&d_shards, (void*)&K, (void*)&hp, (void*)&epochs, (void*)&s0, (void*)&s1, (void*)&ws};
  URE_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kThreads), args, smem, st));
  return 0;
}

}  // namespace
}  // namespace ure

extern "C" int64_t ure_mf_train_workspace_bytes(void) { return (int64_t)sizeof(ure::Workspace); }

extern "C" int ure_mf_train(const ure_mf_shard_t* d_shards, int n_shards, const ure_mf_hparams_t* h_hp,
                            int epochs, int64_t step_begin, int64_t step_end, void* d_workspace,
                            void* stream) {
  using namespace ure;
  URE_REQUIRE(d_shards && h_hp && d_workspace, URE_EINVAL, "ure_mf_train: null argument");
  URE_REQUIRE(n_shards >= 1 && n_shards <= URE_MAX_SHARDS, URE_EINVAL,
              "ure_mf_train: n_shards=%d outside [1,%d]", n_shards, URE_MAX_SHARDS);
  URE_REQUIRE(h_hp->batch > 0 && h_hp->lr_step > 0 && epochs > 0, URE_EINVAL,
              "ure_mf_train: batch/lr_step/epochs must be positive");
  URE_REQUIRE(h_hp->lazy == 0, URE_EUNSUPPORTED, "ure_mf_train: lazy mode not built in this version");
  if (step_end <= step_begin) return 0;
  auto* ws = static_cast<Workspace*>(d_workspace);
  auto st = static_cast<cudaStream_t>(stream);
  switch (h_hp->d) {
    case 8: return launch<8>(d_shards, n_shards, *h_hp, epochs, step_begin, step_end, ws, st);
    case 16: return launch<16>(d_shards, n_shards, *h_hp, epochs, step_begin, step_end, ws, st);
    case 32: return launch<32>(d_shards, n_shards, *h_hp, epochs, step_begin, step_end, ws, st);
    case 64: return launch<64>(d_shards, n_shards, *h_hp, epochs, step_begin, step_end, ws, st);
    case 128: return launch<128>(d_shards, n_shards, *h_hp, epochs, step_begin, step_end, ws, st);
    default:
      set_error("ure_mf_train: d=%d not in {8,16,32,64,128}", h_hp->d);
      return URE_EUNSUPPORTED;
  }
}

extern "C" int ure_mf_flush(const ure_mf_shard_t*, int, const ure_mf_hparams_t* h_hp, int, int64_t, void*) {
  if (h_hp && h_hp->lazy == 0) return 0;   // dense mode: every row is always current
  ure::set_error("ure_mf_flush: lazy mode not built in this version");
  return URE_EUNSUPPORTED;
}

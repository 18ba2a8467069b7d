// Fused Matrix-Factorization training step for ALL shard models in one persistent launch.
//
// Replaces the body of baseTrain (reference method/utils.py:58-98): per batch
//   pred = sum_t P[u,t]*Q[i,t]; e = pred - r; L += e^2;
//   gP[u] += 2e*Q[i]; gQ[i] += 2e*P[u]            (all with PRE-step weights)
// followed by the dense optim.SGD(momentum, weight_decay) update of scratch.py:65-68
// on every row of both tables, for every shard (the K sequential Scratch.train calls
// of method/sisa.py:33-36,86-89 become one launch).
//
// One cooperative launch, one CTA per SM, all steps inside.  The batch-synchronous semantics
// need two grid-wide barriers per step; on this 148-SM part a barrier costs ~2 us of pure
// latency, so each barrier is split into ARRIVE and WAIT and everything that does not depend
// on the other CTAs is issued in between:
//
//   gradients(t)   gather P[u], Q[i] (16-byte L2 loads), shuffle dot product, loss,
//                  red.global.add.v4.f32 scatter of both gradient rows
//   ARRIVE 1       -- overlap: load this thread's sweep operands w, buf (unchanged by the
//                     gradient phase); warp 1 builds the step tables of t+1
//   WAIT 1
//   sweep(t)       dense SGD on every row of every active shard: g += wd*w; buf = mu*buf+g;
//                  w -= lr*buf; g = 0   (in place)
//   ARRIVE 2       -- overlap: permutation index + record fetch of this warp's chunk of t+1
//   WAIT 2
//
// The visiting order is either an explicit permutation (parity runs against the reference) or
// the inline Feistel permutation (feistel.cuh).
#include "common.cuh"
#include "feistel.cuh"

namespace ure {
namespace {

constexpr int kThreads = 1024;
constexpr int KM = URE_MAX_SHARDS;

struct Workspace {
  unsigned barrier[2];     // one monotonic counter per shard group
  long long* trace;        // optional [steps][gridDim.x][6] SM-clock stamps (ure_mf_train_trace), else NULL
  int trace_steps;
  unsigned debug_flags;    // diagnostics only: 1 = skip gradient REDs, 2 = identity visiting order
  unsigned pad[58];
};
static_assert(sizeof(Workspace) == 256, "workspace layout");

struct ShardPtrs {         // 32 bytes: fetched with two LDS.128
  float* P; float* Q; float* gP; float* gQ;
};

// Step tables, double buffered (built for t+1 while step t is still running).
struct StepTables {
  int item_prefix[KM + 1];           // flattened batch positions of the active shards
  long long row_prefix[2 * KM + 1];  // flattened float4 elements of the dense sweep
  int epoch[KM];                     // -1: shard finished
  int start[KM];                     // first position of the batch inside the epoch
  float lr[KM];
  FeistelKeys keys[KM];
};

// largest s with prefix[s] <= x  (prefix[0] = 0, prefix[nseg] = total > x)
template <typename T>
__device__ __forceinline__ int find_segment(const T* prefix, int nseg, T x) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- split grid barrier (monotonic counter; cooperative launch guarantees co-residency)
__device__ __forceinline__ void barrier_arrive(unsigned* counter) {
  __syncthreads();                        // every warp of the CTA has issued its writes / REDs
  if (threadIdx.x == 0) {
    __threadfence();                      // cumulative: orders the CTA's prior writes before the arrival
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void barrier_wait(unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// Gradient work of one 32-interaction warp chunk.  Lane l fetched interaction l (u, it, r, shard s or -1);
// a group of G = D/4 lanes processes its G interactions SUB at a time (2*SUB 16-byte gathers in flight).
// UNIFORM: the whole chunk lies in shard tables `tp`; otherwise each interaction looks its shard up.
template <int D, bool UNIFORM>
__device__ __forceinline__ float chunk_gradients(int u, int it, float r, int s, const ShardPtrs& tp,
                                                  const ShardPtrs* s_ptr, int gl, unsigned dbg) {
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
      if (sq[q] >= 0 && !(dbg & 1u)) {
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

// Per-group pipeline state kept in registers.
struct GroupCtx {
  int ps, pu, pi;          // prefetched record of this lane's interaction in the warp's first chunk
  float pr;
  int seg0;                // preloaded operands of this thread's first sweep element
  size_t off0;
  float4 w0, b0;
  unsigned target;         // running barrier target of the group's counter
  int cur;                 // which of the group's two step tables is current
};

// NG = number of independent shard groups (shard s belongs to group s % NG).  Shards never interact, so
// with NG = 2 the two groups run as two interleaved pipelines: every barrier WAIT of one group is covered
// by gradient / sweep work of the other.
template <int D, int NG>
__global__ void __launch_bounds__(kThreads, 1)
mf_train_kernel(const ure_mf_shard_t* __restrict__ shards, int K, ure_mf_hparams_t hp, int epochs,
                long long step_begin, long long step_end, Workspace* ws) {
  constexpr int G = D / 4;                 // lanes per interaction
  // ---- per-CTA shard tables (static shared memory): constant for the whole launch
  __shared__ ShardPtrs s_ptr[KM];
  __shared__ const ure_inter_t* s_inter[KM];
  __shared__ const int32_t* s_perm[KM];
  __shared__ float* s_buf[2 * KM];         // bufP, bufQ
  __shared__ double* s_sse[KM];
  __shared__ int s_n[KM], s_nuser[KM], s_nitem[KM], s_spe[KM], s_shard_id[KM];
  __shared__ uint32_t s_seed[KM];
  __shared__ FeistelDomain s_dom[KM];
  // ---- per-shard schedule cursor (advanced by the table builder) and the double-buffered step tables
  __shared__ int s_cur_epoch[KM], s_cur_batch[KM];
  extern __shared__ __align__(16) unsigned char dyn_smem[];     // StepTables[NG][2] (static limit is 48 KB)
  StepTables* const s_tab = reinterpret_cast<StepTables*>(dyn_smem);
  __shared__ float s_sse_acc[KM];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int gl = lane % G;
  for (int s = tid; s < K; s += kThreads) {
    const ure_mf_shard_t sh = shards[s];
    s_ptr[s] = ShardPtrs{sh.P, sh.Q, sh.gP, sh.gQ};
    s_inter[s] = sh.inter; s_perm[s] = sh.perm;
    s_buf[2 * s] = sh.bufP; s_buf[2 * s + 1] = sh.bufQ;
    s_sse[s] = sh.sse;
    s_n[s] = sh.n; s_nuser[s] = sh.n_user; s_nitem[s] = sh.n_item;
    const int spe = (sh.n + hp.batch - 1) / hp.batch;
    s_spe[s] = spe;
    s_seed[s] = sh.perm_seed; s_shard_id[s] = sh.shard_id;
    s_dom[s].init((uint32_t)sh.n);
    s_cur_epoch[s] = spe > 0 ? (int)(step_begin / spe) : epochs;     // the only divisions of the launch
    s_cur_batch[s] = spe > 0 ? (int)(step_begin % spe) : 0;
    s_sse_acc[s] = 0.f;
  }
  const unsigned dbg = ws->debug_flags;
  long long* const trace = ws->trace;
  const int trace_steps = ws->trace_steps;
  const long long n_threads = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + tid;
  const int n_warps = (int)(n_threads >> 5);
  const int gwarp = (int)(gtid >> 5);
  const float wd = hp.weight_decay, mu = hp.momentum;
  __syncthreads();

  // Step tables of group g for the step its shards' cursors point at, then advance the cursors.  Executed
  // by ONE warp (lanes = shards, 32 at a time, prefix carried): no block-wide synchronisation inside.
  auto build_tables = [&](StepTables& tb, int g) {
    int carry = 0;
    long long carry2 = 0;
    for (int base = 0; base < K; base += 32) {
      const int s = base + lane;
      int cnt = 0, ep = -1, st = 0;
      long long rows_u = 0, rows_i = 0;
      if (s < K) {
        const int spe = s_spe[s];
        ep = s_cur_epoch[s];
        const int bi = s_cur_batch[s];
        if ((s % NG) == g && spe > 0 && ep < epochs) {
          st = bi * hp.batch;
          cnt = min(hp.batch, s_n[s] - st);
          rows_u = (long long)s_nuser[s] * G;
          rows_i = (long long)s_nitem[s] * G;
          double lr = (double)hp.lr0;
          for (int q = ep / hp.lr_step; q > 0; --q) lr *= (double)hp.lr_decay;
          tb.lr[s] = (float)lr;
          tb.keys[s].init(perm_key(s_seed[s], (uint32_t)s_shard_id[s], (uint32_t)ep));
          if (bi + 1 == spe) { s_cur_epoch[s] = ep + 1; s_cur_batch[s] = 0; }
          else s_cur_batch[s] = bi + 1;
        } else {
          ep = -1;
        }
        tb.epoch[s] = ep;
        tb.start[s] = st;
      }
      int v = cnt;
      long long v2 = rows_u + rows_i;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, v, o);
        const long long b = __shfl_up_sync(0xffffffffu, v2, o);
        if (lane >= o) { v += a; v2 += b; }
      }
      if (s < K) {
        tb.item_prefix[s + 1] = v + carry;
        tb.row_prefix[2 * s + 1] = v2 + carry2 - rows_i;
        tb.row_prefix[2 * s + 2] = v2 + carry2;
      }
      carry += __shfl_sync(0xffffffffu, v, 31);
      carry2 += __shfl_sync(0xffffffffu, v2, 31);
    }
    if (lane == 0) { tb.item_prefix[0] = 0; tb.row_prefix[0] = 0; }
  };

  // Visiting-order index + record of this lane's interaction in warp chunk `wc` of the step in `tb`.
  auto fetch = [&](const StepTables& tb, int wc, int& s, int& u, int& it, float& r) {
    const int item = wc * 32 + lane;
    s = -1; u = 0; it = 0; r = 0.f;
    if (item < tb.item_prefix[K]) {
      s = find_segment(tb.item_prefix, K, item);
      const int j = tb.start[s] + (item - tb.item_prefix[s]);
      const int32_t* pm = s_perm[s];
      uint32_t idx;
      if (pm) idx = (uint32_t)__ldg(pm + (long long)tb.epoch[s] * s_n[s] + j);
      else if (dbg & 2u) idx = (uint32_t)j;
      else idx = feistel(s_dom[s], tb.keys[s], (uint32_t)j);
      const int4 rec = ld_stream_i4(s_inter[s] + idx);
      u = rec.x; it = rec.y; r = __int_as_float(rec.z);
    }
  };

  // ---------------------------------------------------------------- gradients of group g
  auto gradients = [&](GroupCtx& c, int g) {
    const StepTables& tb = s_tab[2 * g + c.cur];
    const int total = tb.item_prefix[K];
    float acc = 0.f;          // sum of e^2 of the chunks this warp processed, all in shard acc_s
    int acc_s = -1;
    for (int wc = gwarp; wc * 32 < total; wc += n_warps) {
      int s, u, it;
      float r;
      if (wc == gwarp) { s = c.ps; u = c.pu; it = c.pi; r = c.pr; }      // prefetched behind a barrier
      else fetch(tb, wc, s, u, it, r);
      const bool valid = s >= 0;
      // the common case: the whole 32-interaction chunk lies in one shard -> table pointers once per warp
      const int s0 = __shfl_sync(0xffffffffu, s, 0);
      const bool uniform = __all_sync(0xffffffffu, s == s0 || !valid);
      const ShardPtrs tp = s_ptr[s0];
      const float my_e = uniform ? chunk_gradients<D, true>(u, it, r, s, tp, s_ptr, gl, dbg)
                                 : chunk_gradients<D, false>(u, it, r, s, tp, s_ptr, gl, dbg);
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
  };

  auto arrive = [&](GroupCtx& c, int g) {
    barrier_arrive(&ws->barrier[g]);
    c.target += gridDim.x;
  };
  auto wait = [&](GroupCtx& c, int g) { barrier_wait(&ws->barrier[g], c.target); };

  // Work that only needs the CTA's own gradient phase: publish the loss, preload the sweep operands the
  // gradient phase did not touch, build the group's tables of the next step (warp 1).
  auto after_gradients = [&](GroupCtx& c, int g, bool more) {
    const StepTables& tb = s_tab[2 * g + c.cur];
    for (int s = tid; s < K; s += kThreads) {
      if ((s % NG) != g) continue;
      const float v = s_sse_acc[s];
      if (v != 0.f) {
        atomicAdd(s_sse[s] + tb.epoch[s], (double)v);
        s_sse_acc[s] = 0.f;
      }
    }
    c.seg0 = 0; c.off0 = 0;
    c.w0 = make_float4(0.f, 0.f, 0.f, 0.f); c.b0 = c.w0;
    if (gtid < tb.row_prefix[2 * K]) {
      c.seg0 = find_segment(tb.row_prefix, 2 * K, gtid);
      c.off0 = (size_t)(gtid - tb.row_prefix[c.seg0]) * 4;
      const ShardPtrs tp = s_ptr[c.seg0 >> 1];
      c.w0 = ld_cg_f4(((c.seg0 & 1) ? tp.Q : tp.P) + c.off0);
      c.b0 = ld_cg_f4(s_buf[c.seg0] + c.off0);
    }
    if (warp == 1 && more) build_tables(s_tab[2 * g + (c.cur ^ 1)], g);
  };

  // ---------------------------------------------------------------- dense SGD sweep of group g
  auto sweep = [&](GroupCtx& c, int g) {
    const StepTables& tb = s_tab[2 * g + c.cur];
    const long long total4 = tb.row_prefix[2 * K];
    for (long long x = gtid; x < total4; x += n_threads) {
      int seg;
      size_t off;
      float4 w, b;
      if (x == gtid) { seg = c.seg0; off = c.off0; w = c.w0; b = c.b0; }
      else {
        seg = find_segment(tb.row_prefix, 2 * K, x);
        off = (size_t)(x - tb.row_prefix[seg]) * 4;
        const ShardPtrs tq = s_ptr[seg >> 1];
        w = ld_cg_f4(((seg & 1) ? tq.Q : tq.P) + off);
        b = ld_cg_f4(s_buf[seg] + off);
      }
      const ShardPtrs tp = s_ptr[seg >> 1];
      float* W = (seg & 1) ? tp.Q : tp.P;
      float* Gr = (seg & 1) ? tp.gQ : tp.gP;
      float* Bf = s_buf[seg];
      const float nlr = -tb.lr[seg >> 1];
      float4 gr = ld_cg_f4(Gr + off);
      // torch SGD: d_p = g + wd*w (fma); buf = buf*mu + d_p; w = w + (-lr)*buf (fma)
      gr.x = fmaf(wd, w.x, gr.x); gr.y = fmaf(wd, w.y, gr.y); gr.z = fmaf(wd, w.z, gr.z); gr.w = fmaf(wd, w.w, gr.w);
      b.x = __fadd_rn(__fmul_rn(b.x, mu), gr.x); b.y = __fadd_rn(__fmul_rn(b.y, mu), gr.y);
      b.z = __fadd_rn(__fmul_rn(b.z, mu), gr.z); b.w = __fadd_rn(__fmul_rn(b.w, mu), gr.w);
      w.x = fmaf(nlr, b.x, w.x); w.y = fmaf(nlr, b.y, w.y); w.z = fmaf(nlr, b.z, w.z); w.w = fmaf(nlr, b.w, w.w);
      st_cg_f4(W + off, w);
      st_cg_f4(Bf + off, b);
      st_cg_f4(Gr + off, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  };

  // visiting-order index + record of this warp's first chunk of the group's next step (static data)
  auto prefetch_next = [&](GroupCtx& c, int g, bool more) {
    if (more) fetch(s_tab[2 * g + (c.cur ^ 1)], gwarp, c.ps, c.pu, c.pi, c.pr);
  };

#define URE_STAMP(PH)                                                                                  \
  if (trace && tid == 0 && (t - step_begin) < trace_steps)                                             \
    trace[((t - step_begin) * gridDim.x + blockIdx.x) * 6 + (PH)] = clock64();

  // ---- prologue: tables and first-chunk records of the first step of every group
  GroupCtx c[NG];
  if (warp == 0) {
#pragma unroll
    for (int g = 0; g < NG; ++g) build_tables(s_tab[2 * g], g);
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    c[g].cur = 0;
    c[g].target = 0;
    fetch(s_tab[2 * g], gwarp, c[g].ps, c[g].pu, c[g].pi, c[g].pr);
  }

  for (long long t = step_begin; t < step_end; ++t) {
    const bool more = t + 1 < step_end;
    URE_STAMP(0)
    if (NG == 1) {
      gradients(c[0], 0);
      URE_STAMP(1)
      arrive(c[0], 0);                     // ARRIVE 1
      after_gradients(c[0], 0, more);
      URE_STAMP(2)
      wait(c[0], 0);                       // WAIT 1
      URE_STAMP(3)
      sweep(c[0], 0);
      URE_STAMP(4)
      arrive(c[0], 0);                     // ARRIVE 2
      prefetch_next(c[0], 0, more);
      wait(c[0], 0);                       // WAIT 2
      URE_STAMP(5)
    } else {
      // two interleaved pipelines; group 1 runs half a step behind group 0
      gradients(c[0], 0);
      arrive(c[0], 0);                     // A: ARRIVE 1
      after_gradients(c[0], 0, more);
      URE_STAMP(1)
      if (t > step_begin) wait(c[NG - 1], NG - 1);          // B: WAIT 2 of step t-1
      gradients(c[NG - 1], NG - 1);
      arrive(c[NG - 1], NG - 1);           // B: ARRIVE 1
      after_gradients(c[NG - 1], NG - 1, more);
      URE_STAMP(2)
      wait(c[0], 0);                       // A: WAIT 1
      sweep(c[0], 0);
      arrive(c[0], 0);                     // A: ARRIVE 2
      prefetch_next(c[0], 0, more);
      URE_STAMP(3)
      wait(c[NG - 1], NG - 1);             // B: WAIT 1
      sweep(c[NG - 1], NG - 1);
      arrive(c[NG - 1], NG - 1);           // B: ARRIVE 2
      prefetch_next(c[NG - 1], NG - 1, more);
      URE_STAMP(4)
      wait(c[0], 0);                       // A: WAIT 2
      URE_STAMP(5)
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) c[g].cur ^= 1;
  }
  if (NG == 2 && step_end > step_begin) wait(c[NG - 1], NG - 1);   // B: WAIT 2 of the last step
#undef URE_STAMP
}

bool g_two_pipelines = false;   // diagnostics (ure_mf_debug_flags bit 2)

template <int D, int NG>
int launch_ng(const ure_mf_shard_t* d_shards, int K, const ure_mf_hparams_t& hp, int epochs, long long s0,
              long long s1, Workspace* ws, cudaStream_t st) {
  auto kern = mf_train_kernel<D, NG>;
  const size_t smem = 2 * NG * sizeof(StepTables);
  URE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  URE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
  URE_REQUIRE(occ >= 1, URE_ECOOP, "mf_train_kernel<%d> cannot be resident", D);
  const int grid = num_sms();
  URE_CUDA(cudaMemsetAsync(ws->barrier, 0, 2 * sizeof(unsigned), st));
  void* args[] = {(void*)&d_shards, (void*)&K, (void*)&hp, (void*)&epochs, (void*)&s0, (void*)&s1, (void*)&ws};
  URE_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kThreads), args, smem, st));
  return 0;
}

// Default: a single pipeline over all shards.  The two-pipeline variant (debug flag 4) is kept for
// experiments: measured 2x SLOWER on B200 (profiles/r1_notes.md) because every ARRIVE's fence and every
// WAIT still stop the whole CTA at a __syncthreads -- it needs a dedicated barrier warp to pay off.
template <int D>
int launch(const ure_mf_shard_t* d_shards, int K, const ure_mf_hparams_t& hp, int epochs, long long s0,
           long long s1, Workspace* ws, cudaStream_t st) {
  if (K >= 2 && g_two_pipelines) return launch_ng<D, 2>(d_shards, K, hp, epochs, s0, s1, ws, st);
  return launch_ng<D, 1>(d_shards, K, hp, epochs, s0, s1, ws, st);
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

// Diagnostics: the next ure_mf_train calls on this workspace record, for their first `steps` steps, six
// SM-clock stamps per CTA (step start, gradients issued, arrive-1 overlap done, barrier 1 passed, sweep
// issued, barrier 2 passed) into d_trace [steps][grid][6] int64.  d_trace = NULL switches tracing off.
extern "C" int ure_mf_train_trace(void* d_workspace, int64_t* d_trace, int steps, void* stream) {
  using namespace ure;
  URE_REQUIRE(d_workspace, URE_EINVAL, "ure_mf_train_trace: null workspace");
  auto* ws = static_cast<Workspace*>(d_workspace);
  long long* p = reinterpret_cast<long long*>(d_trace);
  URE_CUDA(cudaMemcpyAsync(&ws->trace, &p, sizeof(p), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  URE_CUDA(cudaMemcpyAsync(&ws->trace_steps, &steps, sizeof(int), cudaMemcpyHostToDevice,
                           static_cast<cudaStream_t>(stream)));
  URE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int ure_mf_debug_flags(void* d_workspace, unsigned flags, void* stream) {
  using namespace ure;
  URE_REQUIRE(d_workspace, URE_EINVAL, "ure_mf_debug_flags: null workspace");
  auto* ws = static_cast<Workspace*>(d_workspace);
  g_two_pipelines = (flags & 4u) != 0;
  URE_CUDA(cudaMemcpyAsync(&ws->debug_flags, &flags, sizeof(flags), cudaMemcpyHostToDevice,
                           static_cast<cudaStream_t>(stream)));
  URE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int ure_mf_grid_size(void) { return ure::num_sms(); }

extern "C" int ure_mf_flush(const ure_mf_shard_t*, int, const ure_mf_hparams_t* h_hp, int, int64_t, void*) {
  if (h_hp && h_hp->lazy == 0) return 0;   // dense mode: every row is always current
  ure::set_error("ure_mf_flush: lazy mode not built in this version");
  return URE_EUNSUPPORTED;
}

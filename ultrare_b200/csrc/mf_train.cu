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

// dynamic shared memory layout, K = number of shards
struct Smem {
  int K;
  // pointers / constants, loaded once
  const ure_inter_t** inter;
  const int32_t** perm;
  float** P; float** Q; float** bufP; float** bufQ; float** gP; float** gQ;
  double** sse;
  int* n; int* n_user; int* n_item; int* spe;
  uint32_t* seed; int* shard_id;
  // per-step
  int* item_prefix;        // [K+1] flattened batch positions
  long long* row_prefix;   // [2K+1] flattened float4 elements of the dense sweep
  int* epoch; int* start; float* lr;
  Feistel* fe;
  float* sse_acc;          // [K]

  __device__ void carve(unsigned char* base, int K_) {
    K = K_;
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char* p = base + o; o += (bytes + 15) & ~size_t(15); return p; };
    inter = (const ure_inter_t**)take(sizeof(void*) * K);
    perm = (const int32_t**)take(sizeof(void*) * K);
    P = (float**)take(sizeof(void*) * K);
    Q = (float**)take(sizeof(void*) * K);
    bufP = (float**)take(sizeof(void*) * K);
    bufQ = (float**)take(sizeof(void*) * K);
    gP = (float**)take(sizeof(void*) * K);
    gQ = (float**)take(sizeof(void*) * K);
    sse = (double**)take(sizeof(void*) * K);
    row_prefix = (long long*)take(sizeof(long long) * (2 * K + 1));
    n = (int*)take(4 * K); n_user = (int*)take(4 * K); n_item = (int*)take(4 * K); spe = (int*)take(4 * K);
    seed = (uint32_t*)take(4 * K); shard_id = (int*)take(4 * K);
    item_prefix = (int*)take(4 * (K + 1));
    epoch = (int*)take(4 * K); start = (int*)take(4 * K); lr = (float*)take(4 * K);
    fe = (Feistel*)take(sizeof(Feistel) * K);
    sse_acc = (float*)take(4 * K);
  }
  static size_t bytes(int K) {
    auto r = [](size_t b) { return (b + 15) & ~size_t(15); };
    return 9 * r(sizeof(void*) * K) + r(sizeof(long long) * (2 * K + 1)) + 6 * r(4 * K) + r(4 * (K + 1)) +
           3 * r(4 * K) + r(sizeof(Feistel) * K) + r(4 * K);
  }
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

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
mf_train_kernel(const ure_mf_shard_t* __restrict__ shards, int K, ure_mf_hparams_t hp, int epochs,
                long long step_begin, long long step_end, Workspace* ws) {
  constexpr int G = D / 4;                 // lanes per interaction
  constexpr int SUB = (G < 4) ? G : 4;     // interactions whose gathers are in flight together
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Smem sm;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int gl = lane % G;
  if (tid == 0) sm.carve(smem_raw, K);
  __syncthreads();
  for (int s = tid; s < K; s += kThreads) {
    const ure_mf_shard_t sh = shards[s];
    sm.inter[s] = sh.inter; sm.perm[s] = sh.perm;
    sm.P[s] = sh.P; sm.Q[s] = sh.Q; sm.bufP[s] = sh.bufP; sm.bufQ[s] = sh.bufQ;
    sm.gP[s] = sh.gP; sm.gQ[s] = sh.gQ; sm.sse[s] = sh.sse;
    sm.n[s] = sh.n; sm.n_user[s] = sh.n_user; sm.n_item[s] = sh.n_item;
    sm.spe[s] = (sh.n + hp.batch - 1) / hp.batch;
    sm.seed[s] = sh.perm_seed; sm.shard_id[s] = sh.shard_id;
    sm.sse_acc[s] = 0.f;
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
      const int spe = sm.spe[s];
      const bool active = spe > 0 && t < (long long)spe * epochs;
      int ep = 0, cnt = 0, st = 0;
      if (active) {
        ep = (int)(t / spe);
        st = (int)(t % spe) * hp.batch;
        cnt = min(hp.batch, sm.n[s] - st);
        sm.fe[s].init((uint32_t)sm.n[s], perm_key(sm.seed[s], (uint32_t)sm.shard_id[s], (uint32_t)ep));
        sm.lr[s] = (float)((double)hp.lr0 * pow((double)hp.lr_decay, (double)(ep / hp.lr_step)));
      }
      sm.epoch[s] = active ? ep : -1;
      sm.start[s] = st;
      sm.item_prefix[s + 1] = cnt;                               // counts, scanned below
      sm.row_prefix[2 * s + 1] = active ? (long long)sm.n_user[s] * G : 0;
      sm.row_prefix[2 * s + 2] = active ? (long long)sm.n_item[s] * G : 0;
    }
    __syncthreads();
    if (tid < 32) {                                              // warp 0: inclusive scans
      int carry = 0;
      for (int base = 0; base < K; base += 32) {
        int idx = base + lane;
        int v = idx < K ? sm.item_prefix[idx + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (idx < K) sm.item_prefix[idx + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      long long carry2 = 0;
      for (int base = 0; base < 2 * K; base += 32) {
        int idx = base + lane;
        long long v = idx < 2 * K ? sm.row_prefix[idx + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (idx < 2 * K) sm.row_prefix[idx + 1] = v + carry2;
        carry2 += __shfl_sync(0xffffffffu, v, 31);
      }
      if (lane == 0) { sm.item_prefix[0] = 0; sm.row_prefix[0] = 0; }
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase A: gradients
    const int total = sm.item_prefix[K];
    float acc = 0.f;
    int acc_s = -1;
    for (int wc = gwarp; wc * 32 < total; wc += n_warps) {
      const int item = wc * 32 + lane;
      const bool valid = item < total;
      int s = 0;
      int u = 0, it = 0;
      float r = 0.f;
      if (valid) {
        s = find_segment(sm.item_prefix, K, item);
        const int j = sm.start[s] + (item - sm.item_prefix[s]);
        const int32_t* pm = sm.perm[s];
        const uint32_t idx = pm ? (uint32_t)__ldg(pm + (long long)sm.epoch[s] * sm.n[s] + j)
                                : sm.fe[s]((uint32_t)j);
        const int4 rec = ld_stream_i4(sm.inter[s] + idx);
        u = rec.x; it = rec.y; r = __int_as_float(rec.z);
      }
      float my_e = 0.f;
#pragma unroll
      for (int q0 = 0; q0 < G; q0 += SUB) {
        float4 pu[SUB], qi[SUB];
        int uq[SUB], iq[SUB], sq[SUB];
        float rq[SUB];
        bool vq[SUB];
#pragma unroll
        for (int q = 0; q < SUB; ++q) {
          uq[q] = __shfl_sync(0xffffffffu, u, q0 + q, G);
          iq[q] = __shfl_sync(0xffffffffu, it, q0 + q, G);
          rq[q] = __shfl_sync(0xffffffffu, r, q0 + q, G);
          sq[q] = __shfl_sync(0xffffffffu, s, q0 + q, G);
          vq[q] = __shfl_sync(0xffffffffu, (int)valid, q0 + q, G) != 0;
          pu[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          qi[q] = pu[q];
          if (vq[q]) {
            pu[q] = ld_cg_f4(sm.P[sq[q]] + (size_t)uq[q] * D + 4 * gl);
            qi[q] = ld_cg_f4(sm.Q[sq[q]] + (size_t)iq[q] * D + 4 * gl);
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
          if (vq[q]) {
            const float ge = 2.f * e;
            red_add_f4(sm.gP[sq[q]] + (size_t)uq[q] * D + 4 * gl, ge * qi[q].x, ge * qi[q].y, ge * qi[q].z, ge * qi[q].w);
            red_add_f4(sm.gQ[sq[q]] + (size_t)iq[q] * D + 4 * gl, ge * pu[q].x, ge * pu[q].y, ge * pu[q].z, ge * pu[q].w);
          }
        }
      }
      if (valid) {
        if (s != acc_s) {
          if (acc_s >= 0) atomicAdd(&sm.sse_acc[acc_s], acc);
          acc = 0.f;
          acc_s = s;
        }
        acc = fmaf(my_e, my_e, acc);
      }
    }
    if (acc_s >= 0) atomicAdd(&sm.sse_acc[acc_s], acc);
    __syncthreads();
    for (int s = tid; s < K; s += kThreads) {
      const float v = sm.sse_acc[s];
      if (v != 0.f) {
        atomicAdd(sm.sse[s] + sm.epoch[s], (double)v);
        sm.sse_acc[s] = 0.f;
      }
    }
    grid_barrier(&ws->barrier, bar_target);

    // ---------------------------------------------------------------- phase B: dense SGD sweep
    const long long total4 = sm.row_prefix[2 * K];
    const float wd = hp.weight_decay, mu = hp.momentum;
    for (long long x = gtid; x < total4; x += n_threads) {
      const int seg = find_segment(sm.row_prefix, 2 * K, x);
      const size_t off = (size_t)(x - sm.row_prefix[seg]) * 4;
      const int s = seg >> 1;
      float* W = (seg & 1) ? sm.Q[s] : sm.P[s];
      float* Bf = (seg & 1) ? sm.bufQ[s] : sm.bufP[s];
      float* Gr = (seg & 1) ? sm.gQ[s] : sm.gP[s];
      const float nlr = -sm.lr[s];
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
  const size_t smem = Smem::bytes(K);
  URE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

"""Large synthetic shard-training configs of BASELINE.json (C3 / C4 shapes), generated on the GPU.

    python tools/run_config.py --config c3|c3gpu|c4|c4small [--epochs E] [--mode dense|lazy|owner|auto]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_config.py --config c4 ...

Shard s owns users [s*U/K, (s+1)*U/K) (compact user table, local row ids), every shard has the full item
table; shards are spread over the ranks (s mod world) and trained with no communication.  Prints one JSON
line per run with interactions/s (whole job, max over ranks) and the algorithmic-bytes roofline fraction."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrare_b200 import dist as udist, kernels as kn, synth  # noqa: E402

CONFIGS = {
    # name: (n_user, n_item, n_interactions, d, K, batch)
    "c3": (138_493, 26_744, 20_000_263, 64, 8, 30_000),
    "c4": (10_000_000, 1_000_000, 1_000_000_000, 128, 64, 30_000),
    "c4small": (1_250_000, 1_000_000, 125_000_000, 128, 8, 30_000),     # one GPU's share of c4 on an 8-GPU box
    "c3gpu": (17_312, 26_744, 2_500_033, 64, 1, 30_000),                # one GPU's share of c3 on an 8-GPU box
}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3", choices=list(CONFIGS))
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--mode", default="auto", choices=["dense", "lazy", "owner", "auto", "runs"])
ap.add_argument("--max-steps", type=int, default=0, help="train only this many global steps (0 = all)")
a = ap.parse_args()

d_ = udist.init_from_env()
rank, world = d_.rank, d_.world
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
U, I, N, d, K, B = CONFIGS[a.config]
mine = [s for s in range(K) if s % world == rank]
rows_u = U // K
n_shard = N // K
lazy = a.mode == "lazy" or (a.mode == "auto" and (rows_u + I) > 8 * B)
t0 = time.time()
views = kn.alloc_shard_batch([rows_u] * len(mine), I, d, a.epochs, dev, torch.Generator(device=dev).manual_seed(1 + rank),
                             std=0.1)
shards = []
for j, s in enumerate(mine):
    rec = synth.device_interactions(rows_u, I, n_shard, dev, seed=synth.SEED + s)
    P, Q, scratch = views[j]
    shards.append(kn.ShardState(rec, P, Q, a.epochs, shard_id=s + 1, perm_seed=42, scratch=scratch))
sb = kn.ShardBatch(shards, d, B, mode=a.mode if a.mode in ("owner", "runs") else ("lazy" if lazy else "dense"))
torch.cuda.synchronize()
if rank == 0 and sb.owner_plan:
    print("owner plan", sb.owner_plan, file=sys.stderr)
setup_s = time.time() - t0
steps = sb.total_steps if a.max_steps <= 0 else min(sb.total_steps, a.max_steps)
ms_runs, sse_first = [], None
for rep in range(1 if sb.lazy else 2):   # lazy / runs: one launch (their row state is not reset by zeroing buf)   # rep 0 pays the first-launch costs (module load, attribute calls)
    for s_ in shards:
        s_.bufP.zero_(); s_.bufQ.zero_()
        if rep:
            s_.sse.zero_()
    sb.step = 0
    if sb.mode == "owner":
        sb._sched_cover = (0, 0)          # time the schedule pre-pass as well
    d_.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sb.train(steps)
    sb.flush()
    e1.record()
    torch.cuda.synchronize()
    ms_runs.append(d_.max_float(e0.elapsed_time(e1)))
    if rep == 0:
        sse_first = torch.stack([s_.sse for s_ in shards]).sum(0).cpu().numpy()
ms = ms_runs[-1]
inter_local = sum(min(steps * B, s.n * a.epochs) for s in shards)
inter = d_.sum_int(inter_local)
sse = sse_first
if rank == 0:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    alg = inter * (12 + 16 * d)
    print(json.dumps({"config": a.config, "n_gpus": world, "mode": sb.mode, "epochs": a.epochs,
                      "steps": steps, "shards_per_gpu": len(mine), "interactions": inter, "ms": ms,
                      "interactions_per_s": inter / ms * 1e3, "interactions_per_s_per_gpu": inter / ms * 1e3 / world,
                      "algorithmic_GBps_per_gpu": alg / ms / 1e6 / world, "frac_of_hbm_peak": alg / ms / 1e6 / world / peak,
                      "first_epoch_rmse_rank0": float(np.sqrt(sse[0] / max(1, sum(s.n for s in shards)))),
                      "setup_s": setup_s, "mem_GB_rank0": torch.cuda.max_memory_allocated() / 2**30}))

if os.environ.get("URE_TRACE") and rank == 0 and sb.mode in ("dense", "owner"):
    import ctypes as C
    from ultrare_b200 import _lib
    L = _lib.lib()
    grid, nst = L.ure_mf_grid_size(), 40
    trace = torch.zeros((nst, grid, 6), dtype=torch.int64, device=dev)
    _lib.check(L.ure_mf_train_trace(C.c_void_p(sb.ws.data_ptr()), C.c_void_p(trace.data_ptr()), nst, None))
    for s_ in shards:
        s_.bufP.zero_(); s_.bufQ.zero_(); s_.sse.zero_()
    sb.step = 0
    sb.train(min(steps, 60))
    torch.cuda.synchronize()
    tr = trace.cpu().numpy().astype(np.float64)[5:]
    dd = np.diff(tr, axis=2)
    names = ["gradients", "overlap1", "wait1", "sweep", "arr2+fetch+wait2"] if sb.mode == "dense" else \
        ["waves (warp 1)", "other warps", "row sweep", "arrive + next list", "barrier wait"]
    print("per-phase SM cycles (median / p95 / max):")
    for k_, nm in enumerate(names):
        x = dd[:, :, k_].ravel()
        print(f"  {nm:18s} {np.median(x):8.0f} {np.percentile(x, 95):8.0f} {x.max():8.0f}")
    step_cyc = tr[1:, :, 0] - tr[:-1, :, 0]
    print("step cycles median", np.median(step_cyc))

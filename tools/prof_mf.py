"""Profiling driver: one ml1m-shaped K=5 batched MF training (the dominant kernel), nothing else.
    python tools/prof_mf.py [--epochs E] [--d D] [--reps R]
Prints the kernel time measured with CUDA events (never under a profiler)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrare_b200 import kernels as kn, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=30000)
ap.add_argument("--k", type=int, default=5)
ap.add_argument("--mode", default="owner", choices=["dense", "owner"])
ap.add_argument("--trace", action="store_true", help="dense schedule only: per-phase SM-clock stamps")
args = ap.parse_args()

dev = torch.device("cuda:0")
train, _ = synth.ml_like()
u, i, r = train
U, I, d, K = 6040, 3416, 16, args.k
rs = np.random.RandomState(0)
perm_u = rs.permutation(U)
groups = np.array_split(perm_u, K)
row_of = np.zeros(U, dtype=np.int64)
owner = np.zeros(U, dtype=np.int64)
for g, ids in enumerate(groups):
    row_of[ids] = np.arange(len(ids))
    owner[ids] = g
shards = []
gen = torch.Generator(device=dev).manual_seed(1)
for g, ids in enumerate(groups):
    loc = owner[u] == g
    inter = kn.pack_interactions(row_of[u[loc]], i[loc], r[loc] / 5.0, dev)
    P = torch.empty((len(ids), d), device=dev).normal_(generator=gen)
    Q = torch.empty((I, d), device=dev).normal_(generator=gen)
    shards.append(kn.ShardState(inter, P, Q, args.epochs, shard_id=g + 1, perm_seed=42))
sb = kn.ShardBatch(shards, d, args.batch, mode=args.mode)
print("mode", sb.mode, "plan", sb.owner_plan)
n_inter = sum(s.n for s in shards) * args.epochs
for rep in range(args.reps):
    for s in shards:
        s.bufP.zero_(); s.bufQ.zero_(); s.sse.zero_()
    sb.step = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sb.train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {rep}: {ms:.3f} ms for {sb.total_steps} steps ({ms * 1e3 / sb.total_steps:.2f} us/step), "
          f"{n_inter / ms / 1e6:.2f} G inter/s, {n_inter * 268 / ms / 1e6:.1f} GB/s algorithmic")
print("losses", [float(x[-1]) for x in sb.train_losses()])
if sb.mode == "owner":
    import ctypes as C
    from ultrare_b200 import _lib
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().ure_mf_owner_schedule(C.c_void_p(sb.table.data_ptr()), len(sb.shards), C.byref(sb.hp),
                                                    sb.epochs, 0, None))
        e1.record()
        torch.cuda.synchronize()
        print(f"schedule pre-pass ({sb.hp.owner_sched_rows} rows): {e0.elapsed_time(e1):.3f} ms")

if not args.trace:
    sys.exit(0)


def run_trace(flags, label):
    print("====", label)
    L = _lib.lib()
    grid, nst = L.ure_mf_grid_size(), 40
    trace = torch.zeros((nst, grid, 6), dtype=torch.int64, device=dev)
    _lib.check(L.ure_mf_train_trace(C.c_void_p(sb.ws.data_ptr()), C.c_void_p(trace.data_ptr()), nst, None))
    for s in shards:
        s.bufP.zero_(); s.bufQ.zero_(); s.sse.zero_()
    sb.step = 0
    _lib.check(L.ure_mf_debug_flags(C.c_void_p(sb.ws.data_ptr()), flags, None))
    sb.train()
    torch.cuda.synchronize()
    _lib.check(L.ure_mf_train_trace(C.c_void_p(sb.ws.data_ptr()), None, 0, None))
    _lib.check(L.ure_mf_debug_flags(C.c_void_p(sb.ws.data_ptr()), 0, None))
    tr = trace.cpu().numpy().astype(np.float64)[5:]          # skip cold steps
    d = np.diff(tr, axis=2)                                   # [steps, grid, 5]
    names = (["waves (warp 1)", "other warps", "row sweep", "arrive + next list", "barrier wait"] if args.mode == "owner" else
             ["gradients", "overlap1", "wait1", "sweep", "arr2+fetch+wait2"])
    print("per-phase SM cycles (median over CTAs and steps / p95 / max):")
    for k, nm in enumerate(names):
        x = d[:, :, k].ravel()
        print(f"  {nm:18s} {np.median(x):8.0f} {np.percentile(x, 95):8.0f} {x.max():8.0f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s_ in shards:
        s_.bufP.zero_(); s_.bufQ.zero_(); s_.sse.zero_()
    sb.step = 0
    _lib.check(L.ure_mf_debug_flags(C.c_void_p(sb.ws.data_ptr()), flags, None))
    e0.record(); sb.train(); e1.record(); torch.cuda.synchronize()
    _lib.check(L.ure_mf_debug_flags(C.c_void_p(sb.ws.data_ptr()), 0, None))
    ms = e0.elapsed_time(e1)
    print(f"  whole launch {ms:.3f} ms = {ms * 1e3 / sb.total_steps:.2f} us/step, {n_inter * 268 / ms / 1e6:.0f} GB/s algorithmic")
    step_cyc = tr[1:, :, 0] - tr[:-1, :, 0]
    print("step cycles median", np.median(step_cyc), "=> us at 1.965 GHz:", np.median(step_cyc) / 1965)

import ctypes as C
from ultrare_b200 import _lib
if args.mode == "owner":
    cases = ((0, "owner"),)
else:
    cases = ((0, "two warp groups (default split %d)" % sb.warps_group0), (32 << 8, "one group of 32 warps"))
for flags, label in cases:
    run_trace(flags, label)

"""Profiling driver for the OT-grouping kernels (C5 shapes): cost matrix (tcgen05), Sinkhorn, assignment.
    python tools/prof_ot.py [--n N] [--k K] [--d D] [--iters I]
Times each kernel with CUDA events (never under a profiler) and prints achieved GB/s against the
algorithmic bytes of SURVEY.md §8d."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrare_b200 import kernels as kn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--json", type=str, default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
n, k, d = a.n, a.k, a.d
centers = torch.randn((k, d), device=dev, generator=g) * 2.0
X = centers[torch.randint(0, k, (n,), device=dev, generator=g)] + torch.randn((n, d), device=dev, generator=g)
C = X[torch.randperm(n, device=dev, generator=g)[:k]].contiguous()
kp = kn.kpad_for(k)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=a.reps):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


res = {"n": n, "k": k, "d": d, "kpad": kp}
M0 = kn.cost_matrix(X, C)


def cost3():                      # three launches back to back: the host's launch work hides behind the previous one
    for _ in range(3):
        kn.cost_matrix(X, C, out=M0)
    return M0


ms, M = timed(cost3)
ms /= 3
by = 4 * n * d + 4 * n * kp
res["cost_tc"] = {"ms": ms, "GBps": by / ms / 1e6, "TFLOPs": 3 * 2 * n * kp * d / ms / 1e9}
ms_s, Ms = timed(lambda: kn.cost_matrix(X, C, simt=True), reps=2)
res["cost_simt"] = {"ms": ms_s}
res["cost_max_rel_diff"] = float(((M[:, :k] - Ms[:, :k]).abs().max() / Ms[:, :k].abs().max()).item())
scale = float(M[:, :k].min(dim=1).values.mean().item())
sched = [(scale * 0.3, a.iters)]
ms, gpot = timed(lambda: kn.sinkhorn(M, k, sched))
res["sinkhorn_persistent"] = {"ms": ms, "iters": a.iters, "us_per_iter": ms * 1e3 / a.iters,
                              "GBps": (4 * n * kp + 8 * n) * a.iters / ms / 1e6}
colsum = torch.zeros(kp, dtype=torch.float64, device=dev)
g2 = torch.zeros(k, dtype=torch.float32, device=dev)


def split():
    for _ in range(20):
        kn.sinkhorn_colsum(M, k, g2, sched[0][0], n, colsum)
        kn.sinkhorn_update_g(g2, colsum, k, sched[0][0])


ms, _ = timed(split)
res["sinkhorn_split"] = {"ms": ms, "iters": 20, "us_per_iter": ms * 1e3 / 20, "GBps": (4 * n * kp + 8 * n) * 20 / ms / 1e6}
ms, _ = timed(lambda: kn.assign_centroids(M, k, gpot, X))
res["assign_centroids"] = {"ms": ms, "GBps": (4 * n * kp + 4 * n + 4 * n * d) / ms / 1e6}
ms, _ = timed(lambda: kn.sinkhorn_plan(M, k, gpot, sched[0][0]))
res["plan"] = {"ms": ms, "GBps": (4 * n * kp + 4 * n * k) / ms / 1e6}
print(json.dumps(res))
if a.json:
    json.dump(res, open(a.json, "w"), indent=1)

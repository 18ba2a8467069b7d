"""GPU timeline of ONE device-resident retrain-after-delete step of the bench workload (CUPTI via torch.profiler):
every kernel / memcpy / memset with its start offset, duration and the idle gap before it.
    python tools/prof_timeline.py [epochs] [e2e]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultrare_b200 import dist as udist, kernels as kn  # noqa: E402
from ultrare_b200.method.sisa import Sisa  # noqa: E402
from ultrare_b200.read import RatingData, loadData  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 50
E2E = len(sys.argv) > 2 and sys.argv[2] == "e2e"
d = udist.init_from_env()
w = bench.host_workload(0, E)
sp = bench.group_and_split(w, bench.bench_groups(w["n_user"]))
param = bench.Param(w["n_user"], w["n_item"], E)
sp["unlearn_train"] = [kn.pinned_copy(a) for a in sp["unlearn_train"]]
tests = [kn.pinned_copy(t) for t in sp["test"]]
test_np = kn.pinned_copy(np.hstack(tests))
mk = lambda key, sh: [loadData(RatingData(a), bench.BATCH, 1, sh) for a in sp[key]]


def new():
    s = Sisa(param, "mf", 5, sp["group_index"]); s.epoch_eval = "none"; return s


tdl = [loadData(RatingData(t), bench.BATCH, 1, False) for t in tests]
tdata = loadData(RatingData(test_np), bench.BATCH, 1, False)
models = new().learn(mk("learn_train", True), tdl, tdata, 0, "")
train_dl = mk("unlearn_train", True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    if E2E:
        tl = mk("unlearn_train", True)
        tdl2 = [loadData(RatingData(t), bench.BATCH, 1, False) for t in tests]
        td2 = loadData(RatingData(test_np), bench.BATCH, 1, False)
        un = new()
        out = un.unlearn(models, tl, tdl2, td2, list(sp["del_user"]), 0, "")
        kn.download_many([out[0].user_mat.weight.data] + [m.item_mat.weight.data for m in out if getattr(m, "item_mat", None) is not None])
    else:
        un = new()
        un.unlearn(models, train_dl, tdl, tdata, list(sp["del_user"]), 0, "")
    return un


for _ in range(4):
    flush.fill_(1); torch.cuda.synchronize(); step(); torch.cuda.synchronize()
flush.fill_(2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    un = step()
    e1.record()
    torch.cuda.synchronize()
print("event-timed step ms:", e0.elapsed_time(e1), "| host timing:", {k: round(v, 3) if isinstance(v, float) else v for k, v in un.timing.items()})
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = t0
busy = 0.0
print(f"{'start_us':>9} {'dur_us':>9} {'gap_us':>8}  name")
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    gap = s - prev_end
    print(f"{s - t0:9.1f} {en - s:9.1f} {gap:8.1f}  {e.name[:110]}")
    busy += en - s
    prev_end = max(prev_end, en)
print(f"span {prev_end - t0:.1f} us, busy {busy:.1f} us, launches {len(evs)}")

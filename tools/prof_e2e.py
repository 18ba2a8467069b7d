"""Host-side profile of one END-TO-END retrain-after-delete step of the bench workload: host arrays ->
RatingData (pack + upload) -> Sisa.unlearn -> results back on the host (what bench.py's e2e times)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultrare_b200 import dist as udist  # noqa: E402
from ultrare_b200.method.sisa import Sisa  # noqa: E402
from ultrare_b200.read import RatingData, loadData  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 50
d = udist.init_from_env()
w = bench.host_workload(0, E)
rs = np.random.RandomState(1)
perm = rs.permutation(w["n_user"])
sp = bench.group_and_split(w, [perm[i::5].tolist() for i in range(5)])
param = bench.Param(w["n_user"], w["n_item"], E)
mk = lambda key, sh: [loadData(RatingData(a), bench.BATCH, 1, sh) for a in sp[key]]
test_np = np.hstack(sp["test"])


def new():
    s = Sisa(param, "mf", 5, sp["group_index"]); s.epoch_eval = "none"; return s


models = new().learn(mk("learn_train", True), mk("test", False), loadData(RatingData(test_np), bench.BATCH, 1, False), 0, "")
torch.cuda.synchronize()


def e2e(stamps=None):
    t = [time.perf_counter()]
    tl = mk("unlearn_train", True); t.append(time.perf_counter())
    tdl = mk("test", False)
    tdata = loadData(RatingData(test_np), bench.BATCH, 1, False); t.append(time.perf_counter())
    un = new()
    out = un.unlearn(models, tl, tdl, tdata, list(sp["del_user"]), 0, ""); t.append(time.perf_counter())
    merged_h = out[0].user_mat.weight.data.cpu()
    items_h = [m.item_mat.weight.data.cpu() for m in out if getattr(m, "item_mat", None) is not None]
    torch.cuda.synchronize(); t.append(time.perf_counter())
    if stamps is not None:
        stamps.append(np.diff(t) * 1e3)
    return un


for _ in range(3):
    e2e()
st = []
for _ in range(8):
    un = e2e(st)
st = np.array(st)
print("e2e ms: train loaders %.2f | test loaders %.2f | unlearn %.2f | results to host %.2f | total %.2f" %
      (*st.mean(0), st.sum(1).mean()))
print("per-iteration totals", np.round(st.sum(1), 2))
print("unlearn timing", un.timing)
pr = cProfile.Profile()
pr.enable()
e2e()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

"""Host-side profile (cProfile) of one Sisa.unlearn pass of the bench workload."""
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
test_dl = mk("test", False)
test_data = loadData(RatingData(np.hstack(sp["test"])), bench.BATCH, 1, False)
def new():
    s = Sisa(param, "mf", 5, sp["group_index"]); s.epoch_eval = "none"; return s
models = new().learn(mk("learn_train", True), test_dl, test_data, 0, "")
tl = mk("unlearn_train", True)
for _ in range(3):
    new().unlearn(models, tl, test_dl, test_data, list(sp["del_user"]), 0, "")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    new().unlearn(models, tl, test_dl, test_data, list(sp["del_user"]), 0, "")
torch.cuda.synchronize()
print("wall per unlearn: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
un = new()
t0 = time.perf_counter()
un.unlearn(models, tl, test_dl, test_data, list(sp["del_user"]), 0, "")
torch.cuda.synchronize()
print("one unlearn %.3f ms; retrained %s; timing %s" % ((time.perf_counter() - t0) * 1e3, sorted(un.retrain_gid), un.timing))
print("plan sync wait ms", getattr(un._last_batch, "plan_sync_ms", None))
pr = cProfile.Profile()
pr.enable()
new().unlearn(models, tl, test_dl, test_data, list(sp["del_user"]), 0, "")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

"""Diagnostic: the emb-ot grouping on the user table the `--group 0` run saved (synthetic ml1m set)."""
import numpy as np

from ultrare_b200 import synth
from ultrare_b200.config import InsParam, Instance
from ultrare_b200.group import Group
from ultrare_b200.method.utils import ot_cluster_device

synth.ensure_dataset('ml1m')
p = InsParam('ml1m', 2, 24, [64, 32], 5, 2, 'rand')
ins = Instance(p)
X = np.load(ins._user_mat_path(), allow_pickle=True)
shape = type('Shape', (), {'shape': (p.n_user, p.n_item)})()
st = np.random.get_state()
idx = np.random.choice(len(X), size=5, replace=False)
print("init users", idx.tolist())
np.random.set_state(st)
for rep in range(2):
    gi = Group(shape, 'ml1m', X).grouping('ml1m', 5, 'emb-ot', verbose=False, use_cache=False)
    print("Group.grouping sizes", [len(g) for g in gi])
for iters in (1, 2, 10):
    inertia, label, cen, it = ot_cluster_device(X, 5, max_iters=iters, centroid0=X[idx])
    print("same init, outer", iters, "->", it, "inertia", float(inertia), "counts", np.bincount(label, minlength=5).tolist(),
          "nan centroids", int(np.isnan(cen).sum()))

"""Diagnostic: group sizes and per-group train / test rows of the emb-ot grouping on the synthetic ml1m set."""
import sys

import numpy as np

from ultrare_b200 import synth
from ultrare_b200.config import InsParam, Instance
from ultrare_b200.group import Group

synth.ensure_dataset('ml1m')
p = InsParam('ml1m', 2, 24, [64, 32], 5, 2, 'rand')
ins = Instance(p)
user_mat = np.load(ins._user_mat_path(), allow_pickle=True)
shape = type('Shape', (), {'shape': (p.n_user, p.n_item)})()
gi = Group(shape, 'ml1m', user_mat).grouping('ml1m', 5, 'emb-ot', verbose=False)
print("group sizes", [len(g) for g in gi], "distinct users", len(set(sum([list(g) for g in gi], []))))
for host in (False, True):
    import os
    os.environ['URE_HOST_INGEST'] = '1' if host else '0'
    tr, idx, te, tot = ins._read_data(False, 5, gi)
    print("host" if host else "device", "train rows", [len(d) for d in tr], "test rows", [len(d) for d in te], "total", len(tot),
          "group order sizes", [len(g) for g in idx])

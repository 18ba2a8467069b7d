"""Bit-equality check of a schedule pre-pass variant: trains the ml1m-shaped K=5 batch (owner schedule, 3 epochs) and
prints one SHA-256 over every table, momentum array and loss.  Run it once per value of the switch under test
(e.g. URE_SCHED_NT=4 / 2 / -1): the batch lists are identical iff the digests are (no atomics on the path)."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrare_b200 import kernels as kn, synth  # noqa: E402

dev = torch.device("cuda:0")
(u, i, r), _ = synth.ml_like()
U, I, d, K, batch, epochs = 6040, 3416, 16, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 30000, 3
groups = np.array_split(np.random.RandomState(0).permutation(U), K)
row_of, owner = np.zeros(U, dtype=np.int64), np.zeros(U, dtype=np.int64)
for g, ids in enumerate(groups):
    row_of[ids], owner[ids] = np.arange(len(ids)), g
rng = np.random.default_rng(11)
shards = []
for g, ids in enumerate(groups):
    loc = owner[u] == g
    P0 = rng.standard_normal((len(ids), d), dtype=np.float32)
    Q0 = rng.standard_normal((I, d), dtype=np.float32)
    shards.append(kn.ShardState(kn.pack_interactions(row_of[u[loc]], i[loc], (r[loc] / 5.0).astype(np.float32), dev),
                                torch.tensor(P0, device=dev), torch.tensor(Q0, device=dev), epochs, g + 1, 42))
sb = kn.ShardBatch(shards, d, batch, mode="owner")
sb.train()
losses = sb.train_losses()
torch.cuda.synchronize()
h = hashlib.sha256()
for sh in shards:
    for t in (sh.P, sh.Q, sh.bufP, sh.bufQ):
        h.update(t.cpu().numpy().tobytes())
for row in losses:
    h.update(np.asarray(row, dtype=np.float64).tobytes())
print("sched_ab", {k: v for k, v in os.environ.items() if k.startswith("URE_")}, "batch", batch, h.hexdigest()[:24], "loss", float(losses[0][-1]))

"""Time readRating (host filter + split + per-shard upload) against readRatingDevice (one upload + partition kernel)
on an ml1m-sized synthetic table.  python tools/prof_ingest.py [n_rows] [n_group]"""
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, '.')
from ultrare_b200.read import RatingData, readRating, readRatingDevice   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_user, n_item = 6040, 3706
rng = np.random.default_rng(0)
df = pd.DataFrame({0: np.sort(rng.integers(0, n_user, n)), 1: rng.integers(0, n_item, n), 2: rng.integers(1, 6, n)})
del_user = list(rng.choice(n_user, n_user // 50, replace=False))
dev = torch.device('cuda', 0)
torch.zeros(1, device=dev)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host, gi = readRating(df, n_user, 5, del_user, [], K, [], 'a')
    t1 = time.perf_counter()
    ds = [RatingData(h) for h in host]
    RatingData.upload_many(ds, dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dd, gi2, tot = readRatingDevice(df, n_user, 5, del_user, [], K, [], 'a', device=dev)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"iter {it}: host split {1e3*(t1-t0):.1f} ms + upload {1e3*(t2-t1):.1f} ms | device ingest {1e3*(t3-t2):.1f} ms", flush=True)
ok = all(torch.equal(a.records(dev), b.records(dev)) for a, b in zip(ds, dd))
print("equal:", ok)

"""Deterministic synthetic data for the BASELINE configs.

No dataset for the named configs ships with the reference (its data/ml1m holds only the
preprocessing notebook), so "ml1m" runs use a seeded ML-1M-SHAPED file:
6040 users x 3416 items, 896 914 train rows (notebook cell 11), ~10 % per-user test split,
ratings 1..5 with ML-1M's marginal, user-sorted.  Ratings carry a low-rank signal so that
RMSE / NDCG respond to training.  Larger shapes are generated directly on the GPU.
"""
from __future__ import annotations

import os

import numpy as np

SEED = 20231003
ML1M = dict(n_user=6040, n_item=3416, n_train=896914)
RATING_MARGINAL = np.array([0.056, 0.108, 0.261, 0.349, 0.226])      # ratings 1..5 (documented assumption)


def _quantise(score, marginal=RATING_MARGINAL):
    """Map real scores to 1..5 so that the empirical marginal matches `marginal`."""
    cuts = np.quantile(score, np.cumsum(marginal)[:-1])
    return (np.searchsorted(cuts, score, side="right") + 1).astype(np.float64)


def ml_like(n_user=ML1M["n_user"], n_item=ML1M["n_item"], n_train=ML1M["n_train"], seed=SEED, min_per_user=20,
            rank=8):
    """(train, test): each a tuple (users int64, items int64, ratings float64 in 1..5), user-sorted."""
    rng = np.random.default_rng(seed)
    n_total = int(round(n_train / 0.897))
    raw = rng.lognormal(mean=0.0, sigma=1.0, size=n_user)
    cnt = np.maximum(min_per_user, np.round(raw / raw.sum() * n_total)).astype(np.int64)
    cnt = np.minimum(cnt, int(0.6 * n_item))
    n_test_u = np.maximum(1, cnt // 10)
    n_train_u = cnt - n_test_u
    diff = n_train - int(n_train_u.sum())
    order = np.argsort(-cnt)
    j = 0
    while diff != 0:                                      # settle the train total on the heaviest users
        u = order[j % n_user]
        step = 1 if diff > 0 else -1
        if 0 < n_train_u[u] + step and cnt[u] + step <= n_item:
            n_train_u[u] += step
            cnt[u] += step
            diff -= step
        j += 1
    pop = (np.arange(1, n_item + 1, dtype=np.float64)) ** -0.9          # power-law item popularity
    rng.shuffle(pop)
    logp = np.log(pop / pop.sum())
    Pu = rng.standard_normal((n_user, rank))
    Qi = rng.standard_normal((n_item, rank))
    us, its, sc, is_test = [], [], [], []
    for u in range(n_user):
        c = int(cnt[u])
        keys = logp + rng.gumbel(size=n_item)                              # Gumbel top-k = sampling w/o replacement
        items = np.sort(np.argpartition(-keys, c - 1)[:c])
        s = Qi[items] @ Pu[u] / np.sqrt(rank) + 0.7 * rng.standard_normal(c)
        t = np.zeros(c, dtype=bool)
        t[rng.choice(c, int(n_test_u[u]), replace=False)] = True
        us.append(np.full(c, u, dtype=np.int64))
        its.append(items.astype(np.int64))
        sc.append(s)
        is_test.append(t)
    us, its, sc, is_test = map(np.concatenate, (us, its, sc, is_test))
    r = _quantise(sc)
    tr, te = ~is_test, is_test
    return (us[tr], its[tr], r[tr]), (us[te], its[te], r[te])


def write_csv(path, triple):
    u, i, r = triple
    os.makedirs(os.path.dirname(path), exist_ok=True)
    import pandas as pd
    pd.DataFrame({0: u, 1: i, 2: r}).to_csv(path, header=False, index=False)


def ensure_dataset(dataset: str) -> None:
    """Create data/<dataset>/ files under DATA_DIR when they do not exist."""
    from .config import DATASETS
    from .group import DATA_DIR
    tr, te = DATA_DIR + DATASETS[dataset][0], DATA_DIR + DATASETS[dataset][1]
    if os.path.exists(tr) and os.path.exists(te):
        return
    if dataset == "ml1m":
        train, test = ml_like()
    elif dataset == "toy":
        gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "toy_data.npz")
        z = np.load(gold)
        train = (z["train_u"].astype(np.int64), z["train_i"].astype(np.int64), z["train_r2"] / 2.0)
        test = (z["test_u"].astype(np.int64), z["test_i"].astype(np.int64), z["test_r2"] / 2.0)
    else:
        raise ValueError(dataset)
    write_csv(tr, train)
    write_csv(te, test)


def device_interactions(n_user, n_item, n, device, seed=SEED, chunk=1 << 26):
    """int32 [n,4] ure_inter_t records generated ON the GPU (large synthetic shapes):
    users ~ u^1.5 * n_user, items ~ u^2 * n_item (heavy-tailed), ratings from the ML-1M marginal."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rec = torch.empty((n, 4), dtype=torch.int32, device=device)
    cdf = torch.tensor(np.cumsum(RATING_MARGINAL), dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        u = (torch.rand(m, generator=g, device=device) ** 1.5 * n_user).to(torch.int32).clamp_(max=n_user - 1)
        it = (torch.rand(m, generator=g, device=device) ** 2.0 * n_item).to(torch.int32).clamp_(max=n_item - 1)
        r = (torch.bucketize(torch.rand(m, generator=g, device=device), cdf).clamp_(max=4) + 1).to(torch.float32) / 5.0
        rec[lo:hi, 0], rec[lo:hi, 1] = u, it
        rec[lo:hi, 2] = r.view(torch.int32)
        rec[lo:hi, 3] = 0
    return rec

"""Oracle: SISA shard bookkeeping -- deletion set, shard materialisation, routing, merge.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates
  * deletion set                         /root/reference/config.py:46-49 (+ Appendix A1: /100)
  * uniform grouping                     /root/reference/read.py:21-33
  * group re-ordering by rating count    /root/reference/read.py:39-50,73-81,102-106
  * per-group row filter minus deletions /root/reference/read.py:52-68
  * affected-shard routing               /root/reference/method/sisa.py:76-81
  * merge (learn / unlearn)              /root/reference/method/sisa.py:52-58,107-113
"""
from __future__ import annotations

import numpy as np


def deletion_set(n_user, del_per, seed=0):
    """config.py:46-49 with the intent fix n_del = int(del_per/100 * n_user)."""
    rs = np.random.RandomState(seed)
    n_del = int(del_per / 100 * n_user)
    return rs.choice(n_user, n_del, replace=False)


def uniform_groups(n_user, n_group, seed=0):
    """read.py:21-33."""
    group_len = int(np.ceil(n_user / n_group))
    org = np.arange(n_user).tolist()
    if n_group == 1:
        return [org]
    rs = np.random.RandomState(seed)
    rs.shuffle(org)
    return [org[i * group_len:(i + 1) * group_len] for i in range(n_group)]


def sort_groups_by_count(users_col, group_index):
    """read.py:45-50 + sort_group(order='a', var='count') read.py:77-81,102-103."""
    counts = [int(np.isin(users_col, idx).sum()) for idx in group_index]
    order = np.argsort(counts)
    return [group_index[j] for j in order], order


def read_rating(users, items, ratings, n_user, max_rating=5, del_user=(), n_group=1,
                group_index=(), sort="r"):
    """readRating on in-memory columns (read.py:9-70); returns (rating_lists, group_index).

    rating_lists[g] is a float64 [3, n_g] array (uid, iid, rating/max_rating) in file order.
    """
    users = np.asarray(users)
    if len(group_index) == 0:
        group_index = uniform_groups(n_user, n_group)
    group_index = list(group_index)
    if sort in ("d", "a"):
        group_index, _ = sort_groups_by_count(users, group_index)
    del_user = set(int(x) for x in del_user)
    out = []
    for g in range(n_group):
        keep = set(int(x) for x in group_index[g]) - del_user
        loc = np.isin(users, list(keep))
        del_user -= keep
        arr = np.stack([users[loc].astype(np.float64), np.asarray(items)[loc].astype(np.float64),
                        np.asarray(ratings)[loc].astype(np.float64) / max_rating])
        out.append(arr)
    return out, group_index


def route_deletions(group_index, del_user):
    """retrain_gid = {i | exists u in del_user with u in group_index[i]} (sisa.py:76-81)."""
    gid = set()
    sets = [set(int(x) for x in g) for g in group_index]
    for user in del_user:
        for i, s in enumerate(sets):
            if int(user) in s:
                gid.add(i)
                break
    return gid


def merge_learn(P_list, group_index):
    """sisa.py:52-58: merged = 0; merged[G_i] = P_i[G_i]."""
    merged = np.zeros_like(P_list[0])
    for P, g in zip(P_list, group_index):
        g = np.asarray(g, dtype=np.int64)
        merged[g] = P[g]
    return merged


def merge_unlearn(P_before, P_list, group_index, retrain_gid):
    """sisa.py:107-113: clone the pre-unlearn table, overwrite retrained owners' rows."""
    merged = P_before.copy()
    for i in retrain_gid:
        g = np.asarray(group_index[i], dtype=np.int64)
        merged[g] = P_list[i][g]
    return merged

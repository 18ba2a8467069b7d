"""B200 mirror of the reference's read.py: shard materialisation and (device) loaders.

``readRating`` / ``sort_group`` keep the reference's signature and return values
(read.py:9-106) with its defects resolved as SURVEY.md Appendix A lists (A8: no dependence
on 'ml1m' appearing in the path, A9: no debug prints).  ``RatingData`` / ``loadData``
(read.py:108-133) keep the attributes the rest of the code touches (``.dataset.users/
items/ratings``, ``len``), but a loader is a thin descriptor: batches are never
materialised on the host -- the training kernel walks the shard's device-resident records
in the epoch's permutation order (the per-sample ``__getitem__`` + collate of the
reference is its #1 cost).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import kernels as kn


def readRating(dir, n_user, max_rating=5, del_user=[], del_rating=[], n_group=1, group_index=[], sort='r'):
    """reference read.py:9-70.  Returns (rating_lists [n_group] of float64 [3,n_g], group_index)."""
    if len(group_index) == 0:
        group_len = int(np.ceil(n_user / n_group))
        org_index = np.arange(n_user).tolist()
        if n_group == 1:
            group_index = [org_index]
        else:
            np.random.seed(0)
            np.random.shuffle(org_index)
            group_index = [org_index[i * group_len:(i + 1) * group_len] for i in range(n_group)]

    ratings = dir if isinstance(dir, pd.DataFrame) else pd.read_csv(dir, header=None, sep=',')
    users = ratings[0].values
    vals = ratings.values

    if sort in ['d', 'a']:
        sorted_index = sort_group(order='a', group_index=group_index, var='count', ratings0=users)
        group_index = [group_index[i] for i in sorted_index]

    # owner map: one pass instead of K np.in1d scans (a user belongs to the first group listing it)
    owner = np.full(max(n_user, int(users.max()) + 1 if len(users) else 0), -1, dtype=np.int64)
    for g in range(n_group - 1, -1, -1):
        owner[np.asarray(group_index[g], dtype=np.int64)] = g
    deleted = np.zeros(len(owner), dtype=bool)
    if len(del_user):
        deleted[np.asarray(list(del_user), dtype=np.int64)] = True
    row_owner = np.where(deleted[users.astype(np.int64)], -1, owner[users.astype(np.int64)])

    rating_lists = []
    for g in range(n_group):
        ratings_group = vals[row_owner == g].T.astype(np.float64).copy()
        if ratings_group.size == 0:
            ratings_group = np.zeros((3, 0), dtype=np.float64)
        ratings_group[2] /= max_rating
        rating_lists.append(ratings_group)
    return rating_lists, group_index


def sort_group(order='a', group_index=[], var='count', ratings0=[], dataset='ml1'):
    """reference read.py:73-106 (var='count': groups by ascending/descending rating count)."""
    assert var in ['count']
    ratings0 = np.asarray(ratings0).astype(np.int64)
    counts_per_user = np.bincount(ratings0, minlength=int(ratings0.max()) + 1 if len(ratings0) else 1)
    sort_value = []
    for index in group_index:
        idx = np.unique(np.asarray(index, dtype=np.int64))
        idx = idx[idx < len(counts_per_user)]
        sort_value.append(int(counts_per_user[idx].sum()))
    if order == 'a':
        return np.argsort(sort_value)
    return np.argsort(sort_value)[::-1]


class RatingData:
    """reference read.py:108-124: users/items int, ratings float (already / max_rating)."""

    def __init__(self, rating_array):
        self._raw = rating_array              # [3, n]: uid, iid, rating / max_rating (float64, read.py:64-68)
        self._n = int(np.shape(rating_array[0])[0])
        self._cols = {}
        self._records = {}
        self._segments = {}

    def _col(self, j, dtype):
        if j not in self._cols:               # the reference's eager casts (read.py:111-113), done on demand
            self._cols[j] = np.asarray(self._raw[j]).astype(dtype)
        return self._cols[j]

    users = property(lambda self: self._col(0, int))
    items = property(lambda self: self._col(1, int))
    ratings = property(lambda self: self._col(2, float))

    def __len__(self):
        return self._n

    def __getitem__(self, idx):
        return (torch.tensor(self.users[idx], dtype=torch.long),
                torch.tensor(self.items[idx], dtype=torch.long),
                torch.tensor(self.ratings[idx], dtype=torch.float32))

    def records(self, device) -> torch.Tensor:
        """int32 [n,4] ure_inter_t records resident on `device` (uploaded once)."""
        key = str(device)
        if key not in self._records:
            self._records[key] = kn.upload_interactions(self._raw, device)
        return self._records[key]

    def records_mapped(self, device, row_of: torch.Tensor, tag: str) -> torch.Tensor:
        """Records whose user field is row_of[user] (the row inside a compact per-shard user table);
        row_of: int32 tensor on `device`, applied by the pack kernel."""
        key = (str(device), tag)
        if key not in self._records:
            self._records[key] = kn.upload_interactions(self._raw, device, row_of)
        return self._records[key]

    @staticmethod
    def upload_many(datasets, device, row_of=None, tag=None):
        """records()/records_mapped() of several datasets at once: their host copies run in parallel."""
        key = str(device) if tag is None else (str(device), tag)
        todo = [ds for ds in datasets if key not in ds._records]
        for ds, rec in zip(todo, kn.upload_interactions_many([ds._raw for ds in todo], device, row_of)):
            ds._records[key] = rec

    def segments(self, device):
        """(order or None, seg) device tensors of the per-user test segments (utils.py:151-161)."""
        key = str(device)
        if key not in self._segments:
            order, seg = kn.user_segments(self.users)
            self._segments[key] = (None if order is None else kn.upload_array(order, device),
                                   kn.upload_array(seg, device))
        return self._segments[key]


class ShardLoader:
    """What ``loadData`` returns: the DataLoader arguments plus the dataset (read.py:127-133)."""

    def __init__(self, data, batch=30000, n_worker=24, shuffle=True, perms=None):
        self.dataset = data
        self.batch_size = batch
        self.n_worker = n_worker          # accepted for signature parity; there are no host workers
        self.shuffle = shuffle
        self.perms = perms                # optional explicit [epochs][n] visiting orders (parity runs)

    def __len__(self):
        return -(-len(self.dataset) // self.batch_size)

    def explicit_perm(self, device, epochs):
        n = len(self.dataset)
        if self.perms is not None:
            p = np.asarray(self.perms, dtype=np.int32)
            assert p.shape == (epochs, n), f"perms must be [{epochs},{n}]"
            return torch.from_numpy(np.ascontiguousarray(p)).to(device)
        if not self.shuffle:
            return torch.arange(n, dtype=torch.int32, device=device).repeat(epochs, 1)
        return None                       # keyed Feistel permutation inside the kernel


def loadData(data, batch=30000, n_worker=24, shuffle=True, perms=None):
    """reference read.py:127-133."""
    return ShardLoader(data, batch, n_worker, shuffle, perms)


def readSparseMat(dir, n_user, n_item, max_rating=5):
    """reference read.py:136-145; only its shape is used on the 'emb-ot' path (config.py:130)."""
    from scipy.sparse import coo_matrix
    ratings = pd.read_csv(dir, header=None, sep=',')
    row = ratings[0].astype(int).values
    col = ratings[1].astype(int).values
    val = ratings[2].astype(float).values / max_rating
    return coo_matrix((val, (row, col)), shape=(n_user, n_item), dtype=np.float32).tocsr()

"""Oracle: the reference's training path AS SHIPPED -- per-sample Dataset + torch DataLoader + autograd.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the "reference path, shimmed, as shipped"
leg of BASELINE.md §3.1: what `python main.py` of the reference actually executes per epoch, restated because
/root/reference does not exist on the GPU box.

Follows
  * ``RatingData.__getitem__``  /root/reference/read.py:108-124  (three tensors built per SAMPLE)
  * ``loadData``                /root/reference/read.py:127-133  (DataLoader(batch, shuffle, num_workers))
  * ``MF``                      /root/reference/method/utils.py:30-43 (nn.Embedding x2, dot product)
  * ``baseTrain`` (var='nor')   /root/reference/method/utils.py:58-65,82,89-91,108 (loss.item() per batch,
                                zero_grad / backward / step)
  * ``optim.SGD``               /root/reference/method/scratch.py:65-68
It is ~10^4 interactions/s: bench.py runs it on a bounded slice and labels the figure extrapolated.
"""
from __future__ import annotations

import time

import numpy as np


def as_shipped_epoch_slice(u, i, r, n_user, n_item, d=16, batch=30000, n_worker=0, lr=1e-3, wd=0.1, mu=0.9,
                           max_samples=60000, seed=0):
    """Train on the first ``max_samples`` rows of (u, i, r) through Dataset/DataLoader/autograd exactly as the
    reference does; returns (interactions, seconds, train_loss)."""
    import torch
    from torch import nn
    from torch.utils.data import DataLoader, Dataset

    class RatingData(Dataset):                                   # read.py:108-124
        def __init__(self, users, items, ratings):
            self.users, self.items, self.ratings = users.astype(int), items.astype(int), ratings.astype(float)

        def __len__(self):
            return len(self.users)

        def __getitem__(self, idx):
            return (torch.tensor(self.users[idx], dtype=torch.long), torch.tensor(self.items[idx], dtype=torch.long),
                    torch.tensor(self.ratings[idx], dtype=torch.float32))

    class MF(nn.Module):                                         # utils.py:30-43
        def __init__(self):
            super().__init__()
            self.user_mat, self.item_mat = nn.Embedding(n_user, d), nn.Embedding(n_item, d)
            nn.init.normal_(self.user_mat.weight, std=1.0)
            nn.init.normal_(self.item_mat.weight, std=1.0)

        def forward(self, uid, iid):
            return (self.user_mat(uid) * self.item_mat(iid)).sum(1)

    torch.manual_seed(seed)
    m = min(len(u), int(max_samples))
    data = RatingData(np.asarray(u[:m]), np.asarray(i[:m]), np.asarray(r[:m]))
    loader = DataLoader(data, batch_size=batch, shuffle=True, num_workers=n_worker)       # read.py:133
    model = MF()
    loss_fn = nn.MSELoss(reduction='sum')                        # scratch.py:45
    opt = torch.optim.SGD(model.parameters(), lr=lr, weight_decay=wd, momentum=mu)        # scratch.py:65-68
    t0 = time.perf_counter()
    train_loss = 0.0
    for user, item, rating in loader:                            # utils.py:58-91
        loss = loss_fn(model(user, item), rating)
        train_loss += loss.item()
        opt.zero_grad()
        loss.backward()
        opt.step()
    dt = time.perf_counter() - t0
    return m, dt, float(np.sqrt(train_loss / m))

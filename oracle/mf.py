"""Oracle: Matrix-Factorization training exactly as the reference does it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in NumPy fp32 / PyTorch-CPU:
  * ``MF.forward``                      /root/reference/method/utils.py:42-43
  * one ``baseTrain`` epoch (var='nor') /root/reference/method/utils.py:46-111
  * ``optim.SGD(lr, weight_decay=lam, momentum)`` built at
    /root/reference/method/scratch.py:65-68 (dense: every row of both tables gets
    weight decay + momentum every step) and ``StepLR(50, 0.95)`` scratch.py:69,80
  * the per-epoch statistic ``sqrt(sum_batches L / N)`` utils.py:82,108
plus the batch schedule of ``DataLoader(batch, shuffle=True)`` read.py:133
(ceil(N/B) batches, last one partial).  Shuffling itself is NOT restated (the
reference draws it from an unseeded torch generator); both sides
of every parity test consume the same explicit permutation, either injected or
the keyed Feistel permutation defined here and mirrored in
ultrare_b200/csrc/feistel.cuh.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32


# ----------------------------------------------------------------------------
# keyed permutation of [0, n): 4-round Feistel network + cycle walking
# ----------------------------------------------------------------------------
def mix32(x):
    """murmur3 finaliser on uint32 (vectorised)."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def perm_key(seed: int, shard: int, epoch: int) -> int:
    k = int(mix32((seed ^ 0x9E3779B9) & 0xFFFFFFFF))
    k = int(mix32((k + shard * 0x85EBCA77 + 1) & 0xFFFFFFFF))
    k = int(mix32((k ^ ((epoch * 0xC2B2AE3D + 0x27D4EB2F) & 0xFFFFFFFF)) & 0xFFFFFFFF))
    return k


FEISTEL_ROUNDS = 4


def round_hash(x):
    """Round function of the Feistel network (csrc/feistel.cuh round_hash): two multiplies, one xor-shift."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x = (x * 0x9E3779B1) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    return x


def feistel_domain(n: int):
    """(a, b): a = ceil(sqrt(n)), b = ceil(n/a) -- the mixed-radix domain [0,a) x [0,b) >= n."""
    import math
    a = max(1, math.isqrt(n - 1) + 1) if n > 1 else 1
    b = max(1, -(-n // a))
    return a, b


def feistel_perm(n: int, key: int) -> np.ndarray:
    """perm[j] for j in [0,n): the position-j element of the epoch's shuffled order.

    4-round alternating Feistel over x = L*b + R, cycle-walked into [0,n)
    (ultrare_b200/csrc/feistel.cuh)."""
    if n <= 1:
        return np.zeros(n, dtype=np.int64)
    a, b = feistel_domain(n)
    rk = [int(mix32((key + r * 0x9E3779B9) & 0xFFFFFFFF)) for r in range(FEISTEL_ROUNDS)]
    x = np.arange(n, dtype=np.uint64)
    out = np.empty(n, dtype=np.int64)
    todo = np.arange(n)
    while todo.size:
        L = x // b
        R = x - L * b
        for r in range(FEISTEL_ROUNDS):
            if r % 2 == 0:
                L = L + ((round_hash(R ^ rk[r]) * a) >> 32)
                L = np.where(L >= a, L - a, L)
            else:
                R = R + ((round_hash(L ^ rk[r]) * b) >> 32)
                R = np.where(R >= b, R - b, R)
        x = L * b + R
        done = x < n
        out[todo[done]] = x[done].astype(np.int64)
        todo = todo[~done]
        x = x[~done]
    return out


# ----------------------------------------------------------------------------
# fp32 NumPy restatement of one training epoch
# ----------------------------------------------------------------------------
def lr_at_epoch(lr0: float, lr_decay: float, epoch: int, step_size: int = 50) -> float:
    """StepLR(step_size=50, gamma=lr_decay): scratch.py:69,80 (scheduler.step() per epoch)."""
    return lr0 * (lr_decay ** (epoch // step_size))


def sgd_dense(W, buf, g, lr, wd, mu, first):
    """torch.optim.SGD single-tensor update, fp32, dense (scratch.py:65-68).

    d_p = g + wd*W ; buf = d_p (first step) | mu*buf + d_p ; W -= lr*buf
    """
    f = np.float32
    d_p = g + f(wd) * W
    if first:
        buf[...] = d_p
    else:
        buf *= f(mu)
        buf += d_p
    W -= f(lr) * buf


def mf_train_epoch(P, Q, bufP, bufQ, u, i, r, perm, batch, lr, wd, mu, step0):
    """One ``baseTrain`` epoch (utils.py:58-98,108) on fp32 arrays, in place.

    P [U,d], Q [I,d], bufP, bufQ: fp32, modified in place.
    u, i: int arrays [N]; r: fp32 [N] (already rating/max_rating, read.py:66,113);
    perm: visiting order for this epoch; step0: number of optimiser steps already
    taken (0 => momentum buffer initialised from the first gradient).
    Returns (train_loss, sse) with train_loss = sqrt(sse / N) (utils.py:108).
    """
    f = np.float32
    n = len(u)
    sse = 0.0
    step = step0
    for s in range(0, n, batch):
        idx = perm[s:s + batch]
        ub, ib, rb = u[idx], i[idx], r[idx].astype(f)
        pu, qi = P[ub], Q[ib]
        pred = (pu * qi).sum(axis=1, dtype=f)                 # utils.py:43
        e = pred - rb
        sse += float(np.sum(e * e, dtype=f))                   # MSELoss(sum).item(), utils.py:65,82
        ge = (f(2.0) * e)[:, None]
        gP = np.zeros_like(P)
        gQ = np.zeros_like(Q)
        np.add.at(gP, ub, ge * qi)                            # embedding_dense_backward
        np.add.at(gQ, ib, ge * pu)
        sgd_dense(P, bufP, gP, lr, wd, mu, step == 0)         # opt.step(), utils.py:91
        sgd_dense(Q, bufQ, gQ, lr, wd, mu, step == 0)
        step += 1
    return float(np.sqrt(sse / n)), sse


def mf_train(P0, Q0, u, i, r, perms, batch, epochs, lr0=1e-3, wd=0.1, mu=0.9, lr_decay=0.95):
    """``Scratch.train`` numerics without evaluation (scratch.py:65-80)."""
    P, Q = P0.astype(np.float32).copy(), Q0.astype(np.float32).copy()
    bufP, bufQ = np.zeros_like(P), np.zeros_like(Q)
    losses = []
    step = 0
    n = len(u)
    for ep in range(epochs):
        lr = lr_at_epoch(lr0, lr_decay, ep)
        loss, _ = mf_train_epoch(P, Q, bufP, bufQ, u, i, r, perms[ep], batch, lr, wd, mu, step)
        step += -(-n // batch)
        losses.append(loss)
    return P, Q, bufP, bufQ, losses


# ----------------------------------------------------------------------------
# vectorised PyTorch-CPU port of the same arithmetic: the timed CPU arm
# ----------------------------------------------------------------------------
def mf_train_epoch_torch(P, Q, bufP, bufQ, u, i, r, perm, batch, lr, wd, mu, step0):
    """Same recurrence as ``mf_train_epoch`` on torch CPU tensors (all host threads).

    This is the "DataLoader bypassed" form of baseTrain: the
    identical dense SGD-momentum-L2 arithmetic with the per-sample ``__getitem__``
    (read.py:118-124) replaced by index_select.  u, i: int64 tensors; r fp32.
    """
    import torch
    n = u.numel()
    sse = 0.0
    step = step0
    for s in range(0, n, batch):
        idx = perm[s:s + batch]
        ub, ib, rb = u[idx], i[idx], r[idx]
        pu, qi = P[ub], Q[ib]
        e = (pu * qi).sum(1) - rb
        sse += float((e * e).sum())
        ge = (2.0 * e).unsqueeze(1)
        gP = torch.zeros_like(P).index_add_(0, ub, ge * qi)
        gQ = torch.zeros_like(Q).index_add_(0, ib, ge * pu)
        for W, buf, g in ((P, bufP, gP), (Q, bufQ, gQ)):
            g.add_(W, alpha=wd)
            if step == 0:
                buf.copy_(g)
            else:
                buf.mul_(mu).add_(g)
            W.add_(buf, alpha=-lr)
        step += 1
    return float(np.sqrt(sse / n)), sse


# ----------------------------------------------------------------------------
# closed form of the dense update for a row that receives no gradient
# ----------------------------------------------------------------------------
def decay_matrix_power(lr, wd, mu, n):
    """[w;buf] after n gradient-free SGD steps = M^n [w;buf], float64.

    One step with g = 0:  buf' = mu*buf + wd*w ;  w' = w - lr*buf'.
    """
    M = np.array([[1.0 - lr * wd, -lr * mu], [wd, mu]], dtype=np.float64)
    return np.linalg.matrix_power(M, int(n))

"""Oracle: ensemble scoring and RMSE / HR@10 / NDCG@10 as ``baseTest`` computes them.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/method/utils.py:115-210:
  * score  = mean_k (P_k[u] . Q_k[i])              utils.py:141-145
  * rmse   = sqrt(sum (score - r)^2 / n_test)      utils.py:148,163
  * per user (all of the user's test rows, merged across batches utils.py:151-161):
      top_rating = argsort(r)[::-1][:10] ; top_pred = argsort(score)[::-1][:10]
      relevance  = r[top_pred] ; HR = #(relevance >= 4/5) / 10       utils.py:169-176
      common     = in1d(top_rating, top_pred)  (positional!)          utils.py:179
      NDCG       = DCG(relevance*(relevance>=4/5)*common) / DCG(ones(10)),
                   zero padded to length 10                          utils.py:180-181,190-210
  * ndcg, hr = mean over users present in the test data              utils.py:183-184

Tie order: the reference uses NumPy's default (unstable, CPU-dispatch dependent)
argsort; SURVEY.md H7.  ``kind='stable'`` is the documented rule of the CUDA
kernel ("descending value, later index first"); ``kind=None`` reproduces the
reference on *this* host's NumPy.
"""
from __future__ import annotations

import numpy as np

TOP_K = 10
_IDCG = 1.0 + float(np.sum(1.0 / np.log2(np.arange(2, TOP_K + 1))))


def mf_score(P, Q, u, i):
    """MF.forward (utils.py:42-43), fp32."""
    return (P[u] * Q[i]).sum(axis=1, dtype=np.float32)


def ensemble_score(Ps, Qs, u, i):
    """torch.stack(preds).mean(0) over the K models (utils.py:141-145), fp32."""
    preds = np.stack([mf_score(P, Q, u, i) for P, Q in zip(Ps, Qs)]).astype(np.float32)
    return (preds.sum(axis=0, dtype=np.float32) / np.float32(len(Ps))).astype(np.float32)


def sse(score, r):
    d = score.astype(np.float32) - r.astype(np.float32)
    return float(np.sum(d.astype(np.float64) ** 2))


def _dcg(rel):
    return rel[0] + float(np.sum(rel[1:] / np.log2(np.arange(2, len(rel) + 1))))


def user_rank_metrics(r_u, s_u, kind="stable"):
    """HR@10 and NDCG@10 of one user (utils.py:166-181). r_u, s_u float arrays."""
    r_u = np.asarray(r_u, dtype=np.float64)
    s_u = np.asarray(s_u, dtype=np.float64)
    top_rating = np.argsort(r_u, kind=kind)[::-1][:TOP_K]
    top_pred = np.argsort(s_u, kind=kind)[::-1][:TOP_K]
    relevance = r_u[top_pred].copy()
    hit = relevance >= (4 / 5)
    hr = float(hit.sum()) / TOP_K
    common = np.isin(top_rating, top_pred)
    rel = relevance * hit * common
    if len(rel) == 0:
        return hr, 0.0
    rel = np.concatenate([rel, np.zeros(TOP_K - len(rel))])
    return hr, _dcg(rel) / _IDCG


def rank_metrics(u, r, score, kind="stable"):
    """(ndcg, hr) means over the users present in ``u`` (utils.py:166-184).

    Groups by user id in order of first appearance exactly like the host dict
    (utils.py:155-161): a user's rows keep file order even if split over batches.
    """
    u = np.asarray(u)
    order = np.argsort(u, kind="stable")
    us = u[order]
    bounds = np.flatnonzero(np.diff(us)) + 1
    ndcgs, hrs = [], []
    for seg in np.split(order, bounds):
        hr, nd = user_rank_metrics(r[seg], score[seg], kind)
        hrs.append(hr)
        ndcgs.append(nd)
    return float(np.mean(ndcgs)), float(np.mean(hrs))


def base_test(Ps, Qs, u, i, r, kind="stable"):
    """(rmse, ndcg, hr, score) == baseTest(dataloader, models, ...) utils.py:115-187."""
    score = ensemble_score(Ps, Qs, u, i)
    rmse = float(np.sqrt(sse(score, r) / len(u)))
    ndcg, hr = rank_metrics(u, r, score, kind)
    return rmse, ndcg, hr, score

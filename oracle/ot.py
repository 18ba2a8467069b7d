"""Oracle: balanced optimal-transport user grouping (``ot_cluster``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/method/utils.py:628-656:
  centroid0 = X[np.random.choice(n, k, replace=False)]            utils.py:632
  for <=10 iterations:                                             utils.py:635
      dist[j,i] = sum_t (X[i,t]-c[j,t])^2       (fp32, [k,n])      utils.py:637
      inertia   = min_j dist .sum()                                utils.py:638
      trans     = ot.emd(1/n, 1/k, dist.T, 1e-3)                   utils.py:641-644
      label     = argmax_j trans[i,j]           (first max wins)   utils.py:647
      centroid  = per-label mean of X                              utils.py:648
      stop if np.allclose(centroid, new_centroid)                  utils.py:651

``ot.emd`` is POT 0.9.0 (README.md:25), a third-party dependency that is NOT in
/root/reference and cannot be installed offline.  Its published semantics --
exact LP  min <G,M>  s.t. G 1 = a, G^T 1 = b, G >= 0, float64, 4th positional
argument ``numItermax`` (1e-3 -> no cap), returning a basic (vertex) solution --
are restated by ``emd_lp`` with SciPy HiGHS dual simplex.  No golden vector for
it exists in the reference: PARITY UNPINNED for the plan.

``sinkhorn_log`` is the float64 statement of what the CUDA Sinkhorn kernels
compute (north_star subsystem 2); the GPU plan is checked against it at the same
epsilon schedule, iteration count and cost matrix.
"""
from __future__ import annotations

import numpy as np


def cost_matrix(X, C, dtype=np.float64):
    """M[i,j] = sum_t (X[i,t]-C[j,t])^2, [n,k] (utils.py:637 transposed)."""
    X = np.asarray(X, dtype=dtype)
    C = np.asarray(C, dtype=dtype)
    out = np.empty((X.shape[0], C.shape[0]), dtype=dtype)
    for j in range(C.shape[0]):
        d = X - C[j]
        out[:, j] = np.einsum("ij,ij->i", d, d)
    return out


def cost_matrix_ref_fp32(X, C):
    """Exactly the reference expression (fp32 broadcast), returns [k,n] like utils.py:637."""
    X = np.asarray(X, dtype=np.float32)
    C = np.asarray(C, dtype=np.float32)
    return ((X - C[:, np.newaxis]) ** 2).sum(axis=2)


def emd_lp(a, b, M):
    """Exact OT plan [n,k] (float64), vertex solution; stands in for ot.emd (utils.py:644)."""
    from scipy.optimize import linprog
    from scipy.sparse import coo_matrix

    M = np.ascontiguousarray(M, dtype=np.float64)
    n, k = M.shape
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    b = b * (a.sum() / b.sum())                         # POT: b <- b * sum(a)/sum(b)
    nk = n * k
    cols = np.arange(nk)
    rows_r = cols // k                                   # row-sum constraints
    rows_c = n + cols % k                                # column-sum constraints
    A = coo_matrix((np.ones(2 * nk), (np.concatenate([rows_r, rows_c]), np.concatenate([cols, cols]))),
                   shape=(n + k, nk)).tocsr()
    # drop the last (redundant) constraint
    res = linprog(M.ravel(), A_eq=A[:-1], b_eq=np.concatenate([a, b])[:-1], bounds=(0, None),
                  method="highs-ds")
    if res.status != 0:
        raise RuntimeError("emd_lp: " + res.message)
    return res.x.reshape(n, k)


def sinkhorn_log(M, eps_schedule, a=None, b=None, g0=None):
    """Log-domain Sinkhorn on cost M [n,k]; returns (plan, f, g, col_err).

    eps_schedule: iterable of (eps, n_iter).  One iteration, given column
    potentials g:
        f_i   = eps*(log a_i - LSE_j((g_j - M_ij)/eps))          (row scaling)
        P_ij  = exp((f_i + g_j - M_ij)/eps)     (rows sum to a_i exactly)
        g_j  += eps*(log b_j - LSE_i((f_i + g_j - M_ij)/eps))    (column scaling)
    Both scalings are log-sum-exps (the structure of POT's sinkhorn_log): a column whose every entry is hundreds of
    eps below its row's maximum -- stale potentials after the centroids moved -- still has a finite log-sum.
    The returned plan is the row-normalised plan for the final g.
    """
    M = np.asarray(M, dtype=np.float64)
    n, k = M.shape
    loga = np.full(n, -np.log(n)) if a is None else np.log(np.asarray(a, dtype=np.float64))
    logb = np.full(k, -np.log(k)) if b is None else np.log(np.asarray(b, dtype=np.float64))
    g = np.zeros(k) if g0 is None else np.asarray(g0, dtype=np.float64).copy()

    def rows(g, eps):
        T = (g[None, :] - M) / eps
        mx = T.max(axis=1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(T - mx).sum(axis=1))
        return T, lse

    eps = None
    for eps, iters in eps_schedule:
        for _ in range(int(iters)):
            T, lse = rows(g, eps)
            L = T + (loga - lse)[:, None]                    # log P_ij
            cmx = L.max(axis=0)
            g = g + eps * (logb - (cmx + np.log(np.exp(L - cmx[None, :]).sum(axis=0))))
    T, lse = rows(g, eps)
    f = eps * (loga - lse)
    P = np.exp(T + (loga - lse)[:, None])
    col_err = float(np.abs(P.sum(axis=0) - np.exp(logb)).max())
    return P, f, g, col_err


def assign(plan):
    """label = argmax_j plan[i,j], first maximum wins (utils.py:647)."""
    return np.argmax(plan, axis=1)


def balance_labels(M, label, max_aug=1 << 14):
    """Balanced rounding of an assignment (the statement ultrare_b200/csrc/ot_balance.cu is checked against).

    ``label`` is cost-optimal for its own group sizes (the argmax of a Sinkhorn plan row is); users are moved from
    over-full to under-full groups along successive shortest augmenting paths of the k-node group graph with edge
    weights w(j->l) = min_{i in group j} (M[i,l] - M[i,j]) (ties: lowest user index), which keeps optimality and ends
    at the minimum-cost assignment with floor(n/k)..ceil(n/k) users per group -- for k | n and a unique optimum, the
    argmax labels of the exact EMD plan the reference computes (utils.py:644-647).  fp32 arithmetic on the deltas
    and path lengths, Bellman-Ford with (distance, lowest predecessor) ties, exactly as the kernel.
    Returns (label, n_augmentations)."""
    M = np.asarray(M, dtype=np.float32)
    label = np.asarray(label).astype(np.int64).copy()
    n, k = M.shape
    lo, hi = n // k, -(-n // k)
    size = np.bincount(label, minlength=k)
    rows = np.arange(n)
    done = 0
    while done < max_aug:
        if (size > hi).any():
            src, snk = size > hi, size < hi
        elif (size < lo).any():
            src, snk = size > lo, size < lo
        else:
            break
        delta = (M - M[rows, label][:, None]).astype(np.float32)            # [n,k]: M_il - M_i,label(i)
        W = np.full((k, k), np.inf, dtype=np.float32)
        U = np.zeros((k, k), dtype=np.int64)
        for j in range(k):
            mem = np.flatnonzero(label == j)
            if mem.size == 0:
                continue
            dj = delta[mem]
            arg = np.argmin(dj, axis=0)                                        # first minimum = lowest user index
            W[j], U[j] = dj[arg, np.arange(k)], mem[arg]
            W[j, j] = np.inf
        dist = np.where(src, np.float32(0), np.float32(np.inf)).astype(np.float32)
        pred = np.full(k, -1)
        for _ in range(k + 2):
            changed = False
            for j in range(k):
                if not np.isfinite(dist[j]):
                    continue
                for l in range(k):
                    if l == j or not np.isfinite(W[j, l]):
                        continue
                    cand = np.float32(dist[j] + W[j, l])
                    if cand < dist[l] or (cand == dist[l] and pred[l] >= 0 and j < pred[l]):
                        changed |= cand < dist[l]
                        dist[l], pred[l] = cand, j
            if not changed:
                break
        cands = [j for j in range(k) if snk[j] and np.isfinite(dist[j])]
        if not cands:
            break
        sink = min(cands, key=lambda j: (dist[j], j))
        l = sink
        moves = []
        while pred[l] >= 0 and len(moves) <= k:
            j = pred[l]
            moves.append((U[j, l], l))
            l = j
        for u, new in moves:
            label[u] = new
        size[l] -= 1
        size[sink] += 1
        done += 1
    return label, done


def centroid_update(X, label, k):
    """np.array([X[label==i].mean(0) for i in range(k)]) (utils.py:648)."""
    X = np.asarray(X)
    return np.array([X[label == j].mean(axis=0) for j in range(k)])


def ot_cluster(X, k, max_iters=10, rng=None, plan_fn=None, centroid0=None):
    """utils.py:628-656 with a pluggable plan solver (default exact LP).

    rng: object with .choice (defaults to the *global* legacy NumPy RNG, as the
    reference uses, utils.py:632).  Returns (inertia, label, centroid, n_outer).
    """
    X = np.asarray(X)
    n = X.shape[0]
    if centroid0 is None:
        choice = (rng.choice if rng is not None else np.random.choice)
        centroid = X[choice(n, size=k, replace=False)]
    else:
        centroid = np.asarray(centroid0)
    if plan_fn is None:
        plan_fn = lambda a, b, M: emd_lp(a, b, M)
    it = 0
    for it in range(1, max_iters + 1):
        dist = cost_matrix_ref_fp32(X, centroid)                 # [k,n]
        inertia = np.min(dist, axis=0).sum()
        a = np.ones(n) / n
        b = np.ones(k) / k
        trans = plan_fn(a, b, dist.T)
        label = assign(trans)
        new_centroid = centroid_update(X, label, k)
        if np.allclose(centroid, new_centroid):
            break
        centroid = new_centroid
    return inertia, label, centroid, it


def labels_to_groups(label, k):
    """group.py:56-58: list of K ascending user-id lists."""
    label = np.asarray(label)
    return [np.flatnonzero(label == j).tolist() for j in range(k)]

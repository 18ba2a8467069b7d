"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU decomposition (ultrare_b200/dist.py,
SURVEY.md §8e).  The device kernels are stood in for by the oracle's NumPy arithmetic -- what is under test
is the placement, the partial-sum / all-reduce structure and that it reproduces the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ultrare_b200 import dist as udist
    d = udist.init_from_env(backend="gloo")
    try:
        ret[rank] = fn(d)
    finally:
        d.td.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return [ret[r] for r in range(world)]


def _placement(d):
    mine = d.my_shards(range(5))
    lo, hi = d.row_block(6041)
    return dict(rank=d.rank, world=d.world, mine=mine, block=(lo, hi), total=d.sum_int(hi - lo),
                mx=d.max_float(float(d.rank) + 0.5))


def test_placement_and_scalar_collectives():
    out = _run(_placement)
    assert [o["mine"] for o in out] == [[0, 2, 4], [1, 3]]
    assert out[0]["block"] == (0, 3021) and out[1]["block"] == (3021, 6041)
    assert all(o["total"] == 6041 and o["mx"] == 1.5 and o["world"] == 2 for o in out)


def _sinkhorn_sharded(d):
    """Row-sharded Sinkhorn: local column sums + all-reduce of k marginals per iteration (§8e)."""
    from oracle import ot as oot
    rng = np.random.default_rng(0)
    n, k = 2001, 6
    M = rng.random((n, k)) * 20
    sched = [(4.0, 15), (1.0, 25)]
    lo, hi = d.row_block(n)
    Ml = M[lo:hi]
    g = np.zeros(k)
    for eps, iters in sched:
        for _ in range(iters):
            T = (g[None, :] - Ml) / eps
            mx = T.max(1, keepdims=True)
            lse = mx[:, 0] + np.log(np.exp(T - mx).sum(1))
            col = torch.from_numpy(np.exp(T - lse[:, None] - np.log(n)).sum(0))
            d.all_reduce(col)
            g = g + eps * (-np.log(k) - np.log(col.numpy()))
    _, _, g_ref, _ = oot.sinkhorn_log(M, sched)
    # labels + centroid sums: local assignment, all-reduce of [k,d] sums and [k] counts
    lab = np.argmax(g[None, :] - Ml, axis=1)
    cnt = torch.from_numpy(np.bincount(lab, minlength=k).astype(np.int64))
    d.all_reduce(cnt)
    lab_ref = np.argmax(g_ref[None, :] - M, axis=1)
    return dict(err=float(np.abs(g - g_ref).max()), cnt=cnt.numpy().tolist(),
                cnt_ref=np.bincount(lab_ref, minlength=k).tolist())


def test_row_sharded_sinkhorn_equals_single_process():
    out = _run(_sinkhorn_sharded)
    for o in out:
        assert o["err"] < 1e-9 and o["cnt"] == o["cnt_ref"]


def _ensemble_sharded(d):
    """Shard s on rank s mod world: partial ensemble sums all-reduced == single-process ensemble; the merge of
    owner rows as an ALL-GATHER of the ranks' compact tables (padded to the longest rank, Sisa._merged_from) gathered
    back through owner / row maps is bit-exact."""
    from oracle import evalm, sisa as osisa
    rng = np.random.default_rng(1)
    K, U, I, dd, n = 5, 60, 40, 16, 500
    groups = osisa.uniform_groups(U, K)
    Ps = [rng.standard_normal((U, dd), dtype=np.float32) for _ in range(K)]
    Qs = [rng.standard_normal((I, dd), dtype=np.float32) for _ in range(K)]
    u, i = rng.integers(0, U, n), rng.integers(0, I, n)
    merged_ref = osisa.merge_learn(Ps, groups)
    sizes = [len(g) for g in groups]
    by_rank = {r: [s for s in range(K) if d.owner_of_shard(s) == r] for r in range(d.world)}
    maxlen = max(sum(sizes[s] for s in v) for v in by_rank.values())
    send = torch.zeros((maxlen, dd), dtype=torch.float32)
    o = 0
    for s in by_rank[d.rank]:                      # this rank's compact tables (owner rows in group order), back to back
        send[o:o + sizes[s]] = torch.from_numpy(Ps[s][np.asarray(groups[s])])
        o += sizes[s]
    gathered = torch.empty((d.world * maxlen, dd), dtype=torch.float32)
    d.all_gather_into(gathered, send)
    merged = np.zeros_like(merged_ref)
    for r, v in by_rank.items():
        o = r * maxlen
        for s in v:
            merged[np.asarray(groups[s])] = gathered[o:o + sizes[s]].numpy()
            o += sizes[s]
    part = np.zeros(n, dtype=np.float32)
    for s in d.my_shards(range(K)):
        part += evalm.mf_score(merged, Qs[s], u, i)
    tp = torch.from_numpy(part)
    d.all_reduce(tp)
    score = tp.numpy() / np.float32(K)
    ref = evalm.ensemble_score([merged_ref] * K, Qs, u, i)
    return dict(merge_exact=bool(np.array_equal(merged, merged_ref)), err=float(np.abs(score - ref).max()))


def test_sharded_merge_and_ensemble_equal_single_process():
    out = _run(_ensemble_sharded)
    for o in out:
        assert o["merge_exact"] and o["err"] < 1e-5


def _eval_rows_sharded(d):
    """The row-sharded evaluation (Sisa._ensemble_test_sharded): every rank evaluates the test rows of the users its
    shards own against the merged user table and the all-reduced SUM of the item tables; the four sums
    (SSE, NDCG, HR, users) all-reduced equal the single-process ensemble evaluation of the merged test set."""
    from oracle import evalm, sisa as osisa
    rng = np.random.default_rng(4)
    K, U, I, dd, n = 5, 80, 50, 16, 1500
    groups = osisa.uniform_groups(U, K)
    owner = np.zeros(U, dtype=np.int64)
    for s, g in enumerate(groups):
        owner[np.asarray(g)] = s
    merged = rng.standard_normal((U, dd), dtype=np.float32) * 0.3
    Qs = [rng.standard_normal((I, dd), dtype=np.float32) * 0.3 for _ in range(K)]
    u = np.sort(rng.integers(0, U, n))
    i, r = rng.integers(0, I, n), (rng.integers(1, 6, n) / 5).astype(np.float32)
    rmse_ref, ndcg_ref, hr_ref, _ = evalm.base_test([merged] * K, Qs, u, i, r)
    mine = d.my_shards(range(K))
    qsum = torch.from_numpy(np.sum([Qs[s] for s in mine], axis=0, dtype=np.float32) if mine
                            else np.zeros((I, dd), dtype=np.float32))
    d.all_reduce(qsum)
    rows = np.isin(owner[u], mine)                       # the test sets of my shards: the rows of the users I own
    score = evalm.mf_score(merged, qsum.numpy(), u[rows], i[rows]) / np.float32(K)
    users = len(np.unique(u[rows]))
    nd, hr = evalm.rank_metrics(u[rows], r[rows], score) if rows.any() else (0.0, 0.0)
    sums = torch.tensor([evalm.sse(score, r[rows]), nd * users, hr * users, float(users)], dtype=torch.float64)
    d.all_reduce(sums)
    v = sums.numpy()
    return dict(rmse=float(np.sqrt(v[0] / n)), ndcg=float(v[1] / v[3]), hr=float(v[2] / v[3]), users=int(v[3]),
                ref=(rmse_ref, ndcg_ref, hr_ref), users_ref=len(np.unique(u)))


def test_row_sharded_evaluation_equals_single_process():
    out = _run(_eval_rows_sharded)
    for o in out:
        assert o["users"] == o["users_ref"]
        assert abs(o["rmse"] - o["ref"][0]) < 1e-6 and abs(o["ndcg"] - o["ref"][1]) < 1e-6
        assert abs(o["hr"] - o["ref"][2]) < 1e-6


def _grouping_on_two_ranks(d):
    """Instance._group_index under torchrun: only rank 0 may cluster / touch the cache file, and both ranks must
    end with rank 0's grouping (here Group.grouping is stood in by a RANK-DEPENDENT answer)."""
    import tempfile
    from ultrare_b200 import config as cfg
    calls = []

    class FakeGroup:
        def __init__(self, shape, dataset, user_mat):
            pass

        def grouping(self, dataset, n_group, var, verbose=True):
            calls.append(d.rank)
            rs = np.random.RandomState(100 + d.rank)
            perm = rs.permutation(40)
            return [sorted(int(u) for u in perm[i::n_group]) for i in range(n_group)]

    cfg.Group = FakeGroup
    ins = cfg.Instance.__new__(cfg.Instance)
    ins.param = type("P", (), dict(n_user=40, n_item=7, dataset="toy"))()
    ins.timing = {}
    with tempfile.NamedTemporaryFile(suffix=".npy") as f:
        np.save(f.name, np.zeros((40, 4), dtype=np.float32))
        ins._user_mat_path = lambda: f.name
        g = ins._group_index("emb-ot", 4)
        u = ins._group_index("uniform", 4)
    return dict(rank=d.rank, calls=calls, groups=g, uniform=u)


def test_grouping_is_computed_on_rank0_and_identical_on_every_rank():
    out = _run(_grouping_on_two_ranks)
    assert out[0]["calls"] == [0] and out[1]["calls"] == []
    assert out[0]["groups"] == out[1]["groups"] and len(out[0]["groups"]) == 4
    assert sorted(u for g in out[1]["groups"] for u in g) == list(range(40))
    assert out[0]["uniform"] == [] and out[1]["uniform"] == []


def test_mapped_record_cache_is_keyed_on_the_row_map():
    """RatingData.records_mapped must not return the rows of an earlier grouping (stale local rows index
    compact tables out of bounds).  The key carries the map's storage address, version and length."""
    from ultrare_b200 import read
    a = torch.arange(10, dtype=torch.int32)
    b = torch.arange(10, dtype=torch.int32)
    k1, k2 = read._mapped_key("cpu", "t", a), read._mapped_key("cpu", "t", b)
    assert k1 != k2 and k1 == read._mapped_key("cpu", "t", a)
    a[0] = 5                                   # in-place edit bumps the version counter
    assert read._mapped_key("cpu", "t", a) != k1
    ds = read.RatingData(np.zeros((3, 4)))
    ds._records[k1] = "old"
    ds._maps[k1] = a
    ds._drop_mapped("cpu", "t")
    assert k1 not in ds._records and k1 not in ds._maps


def test_ot_cluster_host_loop_with_oracle_kernels(monkeypatch):
    """The host loop of ot_cluster_device (outer iterations, warm start, balanced rounding, centroid update, the
    loud failure on non-finite potentials) with the kernels it calls stood in by the oracle's arithmetic: what is
    under test is the control flow, not the CUDA code (tests/test_gpu_surface.py, tests/test_gpu_kernels.py)."""
    from oracle import ot as oot
    from ultrare_b200 import kernels as kn
    from ultrare_b200.method import utils as mu
    monkeypatch.setattr(mu, "_cuda_device", lambda device=None: torch.device("cpu"))

    def cost_matrix(Xd, Cd, want_inertia=False):
        M = ((Xd[:, None, :].double() - Cd[None, :, :].double()) ** 2).sum(-1).float()
        return M, M.min(1).values.double().sum().reshape(1)

    calls = {"warm": 0, "cold": 0, "balance": 0, "poison": False}

    def sinkhorn(M, k, sched, g=None, tol=0.0):
        _, _, gg, _ = oot.sinkhorn_log(M.numpy()[:, :k], sched, g0=None if g is None else g.numpy())
        out = torch.from_numpy(gg.astype(np.float32))
        calls["cold" if g is None else "warm"] += 1
        if calls["poison"]:
            out[0] = float("nan")
        return out

    def assign_centroids(M, k, g, X):
        lab = torch.argmax(g[None, :k] - M[:, :k], dim=1).to(torch.int32)
        return lab, None, torch.bincount(lab.long(), minlength=k)

    def balance_labels(M, k, label, cnt, max_aug):
        lab, n_aug = oot.balance_labels(M.numpy()[:, :k], label.numpy())
        label.copy_(torch.from_numpy(lab.astype(np.int32)))
        cnt.copy_(torch.bincount(label.long(), minlength=k))
        calls["balance"] += 1
        return torch.tensor([n_aug, 0, 0], dtype=torch.int32)

    def centroid_sums(X, label, k):
        sums = torch.zeros((k, X.shape[1]), dtype=torch.float64)
        sums.index_add_(0, label.long(), X.double())
        return sums, torch.bincount(label.long(), minlength=k)

    for name, fn in dict(cost_matrix=cost_matrix, sinkhorn=sinkhorn, assign_centroids=assign_centroids,
                         balance_labels=balance_labels, centroid_sums=centroid_sums,
                         upload_table=lambda a, dev: torch.from_numpy(np.ascontiguousarray(a)),
                         upload_array=lambda a, dev: torch.from_numpy(np.ascontiguousarray(a)),
                         download_many=lambda ts: [t.numpy().copy() for t in ts]).items():
        monkeypatch.setattr(kn, name, fn)
    rng = np.random.default_rng(2)
    n, k = 602, 4                                       # k does not divide n: sizes floor / ceil
    X = rng.standard_normal((n, 8)).astype(np.float32)
    inertia, label, cen, it = mu.ot_cluster_device(X, k, max_iters=3, centroid0=X[:k].copy())
    cnt = np.bincount(label, minlength=k)
    assert sorted(cnt.tolist()) == [150, 150, 151, 151], cnt
    assert np.isfinite(cen).all() and np.isfinite(inertia) and label.dtype == np.int64
    assert calls["cold"] == 1 and calls["warm"] == it - 1 and calls["balance"] == it
    calls["poison"] = True                              # a NaN potential must not end as a silent degenerate grouping
    with pytest.raises(RuntimeError, match="not finite"):
        mu.ot_cluster_device(X, k, max_iters=2, centroid0=X[:k].copy())


def test_balanced_rounding_oracle_reaches_the_exact_emd_assignment():
    """oracle.ot.balance_labels (what csrc/ot_balance.cu is checked against): from the argmax of a Sinkhorn plan the
    successive-shortest-path rounding ends at the argmax labels of the exact LP plan (the reference's ot.emd,
    utils.py:644-647): equal labels, equal cost, exactly n/k users per group."""
    from oracle import ot as oot
    rng = np.random.default_rng(5)
    n, k = 900, 5
    X = rng.standard_normal((n, 8)).astype(np.float32)
    M = oot.cost_matrix_ref_fp32(X, X[:k]).T.copy()
    scale = float(M.min(1).mean())
    _, _, g, _ = oot.sinkhorn_log(M, [(scale, 10), (0.3 * scale, 20), (0.1 * scale, 30), (0.03 * scale, 60)])
    lab = np.argmax(g[None, :] - M, axis=1)
    assert np.bincount(lab, minlength=k).tolist() != [n // k] * k          # the entropic argmax is not balanced
    bal, n_aug = oot.balance_labels(M, lab)
    assert np.bincount(bal, minlength=k).tolist() == [n // k] * k and n_aug > 0
    emd = oot.assign(oot.emd_lp(np.ones(n) / n, np.ones(k) / k, M))
    assert np.array_equal(bal, emd)
    rows = np.arange(n)
    assert abs(float(M[rows, bal].sum()) - float(M[rows, emd].sum())) < 1e-3


def _talker(rank, q, fail):
    import time
    if fail and rank == 1:
        raise RuntimeError("rank 1 breaks")
    if fail:
        time.sleep(120)                    # rank 0 would wait in a collective
    q.put((rank, rank * 10))


def test_collect_results_ends_at_once_when_a_worker_dies():
    """The harness of the 2-GPU test (conftest.collect_results): results of all workers; and a worker that raises ends
    the wait within seconds -- the survivor is killed -- instead of after the time-out (that cost the round its last
    GPU minutes once)."""
    import time
    from conftest import collect_results
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_talker, args=(r, q, False)) for r in range(2)]
    for p in procs:
        p.start()
    assert collect_results(procs, q, 2, timeout=60) == {0: 0, 1: 10}
    for p in procs:
        p.join(30)
    q = ctx.Queue()
    procs = [ctx.Process(target=_talker, args=(r, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    t0 = time.monotonic()
    with pytest.raises(pytest.fail.Exception, match="exit codes"):
        collect_results(procs, q, 2, timeout=100)
    assert time.monotonic() - t0 < 60 and not any(p.is_alive() for p in procs)

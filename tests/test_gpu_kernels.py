"""GPU parity: every kernel of libultrare_b200.so, called through the C ABI, against the
CPU oracle and against the fixtures the reference itself produced (tests/golden/)."""
import numpy as np
import pytest

from conftest import init_weights, load_gold
from oracle import evalm, mf as omf, ot as oot, sisa as osisa

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _shard(K, toy, dev, epochs, batch, explicit_perm, seeds, shard_ids=None, data=None):
    torch = _torch()
    from ultrare_b200 import kernels as kn
    shards, host = [], []
    for s in range(K):
        u, i, r = data[s] if data is not None else toy["train"]
        r32 = (np.asarray(r, dtype=np.float64) / 5.0).astype(np.float32) if data is None else r
        P0, Q0 = init_weights(seeds[s], toy["n_user"], toy["n_item"], toy["k"])
        sid = s if shard_ids is None else shard_ids[s]
        perms = [omf.feistel_perm(len(u), omf.perm_key(42, sid, ep)) for ep in range(epochs)]
        inter = kn.pack_interactions(u, i, r32.astype(np.float64), dev)
        perm_t = torch.tensor(np.stack(perms).astype(np.int32), device=dev) if explicit_perm else None
        shards.append(kn.ShardState(inter, torch.tensor(P0, device=dev), torch.tensor(Q0, device=dev), epochs,
                                    shard_id=sid, perm_seed=42, perm=perm_t))
        host.append((u, i, r32, P0, Q0, perms))
    return shards, host


MODES = ["dense", "owner"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("explicit_perm", [True, False])
def test_mf_train_vs_reference_golden(toy, cuda_dev, explicit_perm, mode):
    """ure_mf_train == reference baseTrain (+SGD/StepLR): losses 1e-3 rel, weights 1e-4 abs (Appendix E)."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    z = load_gold("toy_train.npz")
    epochs, batch = int(z["epochs"]), int(z["batch"])
    shards, _ = _shard(1, toy, cuda_dev, epochs, batch, explicit_perm, [int(z["weight_seed"])])
    sb = kn.ShardBatch(shards, toy["k"], batch, mode=mode)
    assert sb.mode == mode
    sb.train()
    torch.cuda.synchronize()
    losses = sb.train_losses()[0]
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-3)
    assert np.abs(losses - z["losses"]).max() / z["losses"].max() < 1e-5
    assert np.abs(shards[0].P.cpu().numpy() - z["P_final"]).max() < 1e-4
    assert np.abs(shards[0].Q.cpu().numpy() - z["Q_final"]).max() < 1e-4
    assert float(shards[0].gP.abs().max()) == 0.0 and float(shards[0].gQ.abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["dense", "owner", "lazy"])
def test_mf_train_across_two_steplr_boundaries_vs_reference_golden(toy, cuda_dev, mode):
    """101 epochs in one persistent launch: the kernels' lr_of(epoch) (StepLR(50, 0.95), scratch.py:69,80) crosses
    two decay boundaries; fixture = the reference's own baseTrain + scheduler (tests/golden/toy_steplr.npz).  The
    lazy schedule's closed-form catch-up needs a constant learning rate: it refuses such a run loudly."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    z = load_gold("toy_steplr.npz")
    n, epochs, batch = int(z["n"]), int(z["epochs"]), int(z["batch"])
    u, i, r = (a[:n] for a in toy["train"])
    P0, Q0 = init_weights(int(z["weight_seed"]), int(z["n_user"]), int(z["n_item"]))
    st = kn.ShardState(kn.pack_interactions(u, i, r / 5.0, cuda_dev), torch.tensor(P0, device=cuda_dev),
                       torch.tensor(Q0, device=cuda_dev), epochs, shard_id=int(z["shard_id"]), perm_seed=int(z["perm_seed"]))
    sb = kn.ShardBatch([st], 16, batch, mode=mode)
    assert sb.mode == mode
    if mode == "lazy":
        with pytest.raises(RuntimeError, match="constant"):
            sb.train()
        return
    sb.train()
    losses = sb.train_losses()[0]
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-3)
    assert np.abs(losses - z["losses"]).max() / z["losses"].max() < 2e-5
    assert np.abs(st.P.cpu().numpy() - z["P_final"]).max() < 1e-4
    assert np.abs(st.Q.cpu().numpy() - z["Q_final"]).max() < 1e-4


@pytest.mark.parametrize("mode", MODES)
def test_mf_train_epoch_by_epoch_equals_single_launch(toy, cuda_dev, mode):
    torch = _torch()
    from ultrare_b200 import kernels as kn
    a, _ = _shard(1, toy, cuda_dev, 2, 3000, False, [7])
    b, _ = _shard(1, toy, cuda_dev, 2, 3000, False, [7])
    sa, sb = kn.ShardBatch(a, 16, 3000, mode=mode), kn.ShardBatch(b, 16, 3000, mode=mode)
    sa.train()
    spe = b[0].steps_per_epoch(3000)
    for t in range(1, 2 * spe + 1):
        sb.train(t)
    torch.cuda.synchronize()
    # identical schedule; only the order of fp32 atomics differs
    assert np.abs(a[0].P.cpu().numpy() - b[0].P.cpu().numpy()).max() < 2e-5
    np.testing.assert_allclose(sa.train_losses()[0], sb.train_losses()[0], rtol=1e-6)


@pytest.mark.parametrize("mode", MODES)
def test_mf_train_k_shards_batched_vs_oracle(toy, cuda_dev, mode):
    """K ragged shards in ONE launch (different sizes => different steps/epoch) == K oracle trainings."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    u, i, r = toy["train"]
    r32 = (r / 5.0).astype(np.float32)
    K, epochs, batch = 3, 2, 3000
    groups = osisa.uniform_groups(toy["n_user"], K)
    data = []
    for s, g in enumerate(groups):
        loc = np.isin(u, g if s else g[: len(g) // 2])        # shard 0 much smaller: ragged schedules
        data.append((u[loc], i[loc], r32[loc]))
    shards, host = _shard(K, toy, cuda_dev, epochs, batch, False, [11, 12, 13], data=data)
    assert len({sh.steps_per_epoch(batch) for sh in shards}) > 1
    sb = kn.ShardBatch(shards, toy["k"], batch, mode=mode)
    sb.train()
    torch.cuda.synchronize()
    losses = sb.train_losses()
    for s in range(K):
        uu, ii, rr, P0, Q0, perms = host[s]
        P, Q, _, _, ls = omf.mf_train(P0, Q0, uu, ii, rr, perms, batch, epochs)
        np.testing.assert_allclose(losses[s], ls, rtol=1e-5)
        assert np.abs(shards[s].P.cpu().numpy() - P).max() < 1e-4
        assert np.abs(shards[s].Q.cpu().numpy() - Q).max() < 1e-4


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("d", [8, 32, 64, 128])
def test_mf_train_other_dims_vs_oracle(cuda_dev, d, mode):
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(d)
    U, I, n, batch, epochs = 300, 200, 5000, 1024, 2
    u, i = rng.integers(0, U, n), rng.integers(0, I, n)
    r = rng.integers(1, 6, n).astype(np.float32) / 5
    P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
    Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
    perms = [omf.feistel_perm(n, omf.perm_key(9, 0, ep)) for ep in range(epochs)]
    sh = kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                       torch.tensor(Q0, device=cuda_dev), epochs, 0, 9)
    sb = kn.ShardBatch([sh], d, batch, mode=mode)
    sb.train()
    P, Q, _, _, ls = omf.mf_train(P0, Q0, u, i, r, perms, batch, epochs)
    np.testing.assert_allclose(sb.train_losses()[0], ls, rtol=1e-5)
    assert np.abs(sh.P.cpu().numpy() - P).max() < 1e-4 and np.abs(sh.Q.cpu().numpy() - Q).max() < 1e-4


def test_ensemble_score_and_metrics_vs_reference_golden(toy, cuda_dev):
    """Scores 1e-5 abs, RMSE/HR 1e-3 rel vs the reference's baseTest; NDCG vs the stable-rule oracle."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    z = load_gold("toy_train.npz")
    u, i, r = toy["test"]
    r32 = (r / 5.0).astype(np.float32)
    inter = kn.pack_interactions(u, i, r / 5.0, cuda_dev)
    P, Q = torch.tensor(z["P_final"], device=cuda_dev), torch.tensor(z["Q_final"], device=cuda_dev)
    score, sse = kn.ensemble_score([P], [Q], inter)
    assert np.abs(score.cpu().numpy() - z["test_score"]).max() < 1e-5
    rmse = float(np.sqrt(sse.item() / len(u)))
    assert abs(rmse - float(z["test_rmse"])) / float(z["test_rmse"]) < 1e-5
    order, seg = kn.user_segments(u)
    assert order is None
    out = kn.rank_metrics(inter, score, torch.tensor(seg, device=cuda_dev)).cpu().numpy()
    n_users = len(np.unique(u))
    assert out[2] == n_users
    nd_o, hr_o = evalm.rank_metrics(u, r32, score.cpu().numpy(), kind="stable")
    assert abs(out[1] / n_users - float(z["test_hr"])) / float(z["test_hr"]) < 1e-3
    assert abs(out[1] / n_users - hr_o) < 1e-12
    assert abs(out[0] / n_users - nd_o) < 1e-9
    # GPU scores through the reference's own host ranking (default argsort on this box): H7
    nd_ref_rule, _ = evalm.rank_metrics(u, r32, score.cpu().numpy(), kind=None)
    assert abs(nd_ref_rule - float(z["test_ndcg_ref"])) / float(z["test_ndcg_ref"]) < 1e-3


def test_ensemble_k_models_and_unsorted_users(cuda_dev):
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(0)
    U, I, d, n, K = 50, 40, 16, 3000, 5
    u, i = rng.integers(0, U, n), rng.integers(0, I, n)      # users NOT contiguous -> order indirection
    r = rng.integers(1, 11, n).astype(np.float32) / 10
    Ps = [rng.standard_normal((U, d), dtype=np.float32) for _ in range(K)]
    Qs = [rng.standard_normal((I, d), dtype=np.float32) for _ in range(K)]
    inter = kn.pack_interactions(u, i, r, cuda_dev)
    tP = [torch.tensor(p, device=cuda_dev) for p in Ps]
    tQ = [torch.tensor(q, device=cuda_dev) for q in Qs]
    score, sse = kn.ensemble_score(tP, tQ, inter)
    ref = evalm.ensemble_score(Ps, Qs, u, i)
    assert np.abs(score.cpu().numpy() - ref).max() < 2e-5
    assert abs(sse.item() - evalm.sse(ref, r)) / evalm.sse(ref, r) < 1e-5
    # partial sums + finalize == one call (the multi-GPU evaluation path)
    s1, _ = kn.ensemble_score(tP[:2], tQ[:2], inter, denom=1.0)
    s2, _ = kn.ensemble_score(tP[2:], tQ[2:], inter, denom=1.0)
    fin, sse2 = kn.score_finalize(s1 + s2, inter, float(K))
    assert np.abs(fin.cpu().numpy() - ref).max() < 2e-5 and abs(sse2.item() - sse.item()) / sse.item() < 1e-5
    order, seg = kn.user_segments(u)
    assert order is not None
    out = kn.rank_metrics(inter, score, torch.tensor(seg, device=cuda_dev), torch.tensor(order, device=cuda_dev))
    nd, hr = evalm.rank_metrics(u, r, score.cpu().numpy(), kind="stable")
    out = out.cpu().numpy()
    assert out[2] == len(np.unique(u))
    assert abs(out[0] / out[2] - nd) < 1e-9 and abs(out[1] / out[2] - hr) < 1e-12


def test_rank_metrics_edge_cases(cuda_dev):
    """Users with 1 item, < 10 items (zero padding), all-tied ratings and tied scores."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    u = np.array([0] + [1] * 3 + [2] * 10 + [3] * 25 + [4] * 12)
    n = len(u)
    rng = np.random.default_rng(1)
    r = (rng.integers(1, 6, n) / 5).astype(np.float32)
    r[u == 4] = 0.8
    s = rng.standard_normal(n).astype(np.float32)
    s[u == 3] = np.round(s[u == 3])                        # many tied scores
    inter = kn.pack_interactions(u, np.zeros(n, dtype=np.int64), r, cuda_dev)
    _, seg = kn.user_segments(u)
    out = kn.rank_metrics(inter, torch.tensor(s, device=cuda_dev), torch.tensor(seg, device=cuda_dev)).cpu().numpy()
    nd, hr = evalm.rank_metrics(u, r, s, kind="stable")
    assert out[2] == 5 and abs(out[0] / 5 - nd) < 1e-12 and abs(out[1] / 5 - hr) < 1e-12
    empty = kn.rank_metrics(inter[:0], torch.zeros(0, device=cuda_dev), torch.zeros(1, dtype=torch.int64, device=cuda_dev))
    assert empty.cpu().numpy().tolist() == [0.0, 0.0, 0.0]


def test_eval_jobs_equal_single_calls(cuda_dev):
    """ure_eval_jobs (many evaluations, grid.y = job) = ure_ensemble_score + ure_rank_metrics per job: different model
    counts, test sets of different sizes (one empty-segment user, one shuffled set with the order indirection)."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(11)
    U, I, d = 70, 50, 16
    tabs = [(torch.tensor(rng.standard_normal((U, d), dtype=np.float32), device=cuda_dev),
             torch.tensor(rng.standard_normal((I, d), dtype=np.float32), device=cuda_dev)) for _ in range(4)]
    sets = []
    for n, shuffle in ((4000, False), (900, True), (37, False)):
        u = np.sort(rng.integers(0, U - 1, n))                       # user U-1 has no test rows
        if shuffle:
            u = rng.permutation(u)
        i = rng.integers(0, I, n)
        r = (rng.integers(1, 6, n) / 5).astype(np.float32)
        inter = kn.pack_interactions(u, i, r, cuda_dev)
        order, seg = kn.user_segments(u)
        sets.append((inter, None if order is None else torch.tensor(order, device=cuda_dev), torch.tensor(seg, device=cuda_dev)))
    jobs, want = [], []
    for ms in ([0], [0, 1, 2], [3, 1], [2]):
        for inter, order, seg in sets:
            Ps, Qs = [tabs[m][0] for m in ms], [tabs[m][1] for m in ms]
            jobs.append((Ps, Qs, inter, order, seg))
            score, sse = kn.ensemble_score(Ps, Qs, inter)
            want.append(torch.cat([sse, kn.rank_metrics(inter, score, seg, order)]).cpu().numpy())
    got = kn.eval_jobs(jobs, d).cpu().numpy()
    want = np.stack(want)
    assert got.shape == want.shape == (12, 4)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_rank_metrics_long_segments(cuda_dev):
    """Segments far longer than a warp (several elements per lane in the staged arg-max rounds) and longer than the
    shared-memory stage (the counting-rank path), with heavily tied scores and ratings, in user order and through
    the order indirection."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(3)
    lens = [300, 513, 700, 33, 64, 9, 512]
    u = np.concatenate([np.full(n, j) for j, n in enumerate(lens)])
    n = len(u)
    r = (rng.integers(1, 6, n) / 5).astype(np.float32)
    s = np.round(rng.standard_normal(n) * 2).astype(np.float32) / 2          # many exact ties
    for shuffle in (False, True):
        idx = rng.permutation(n) if shuffle else np.arange(n)
        uu, rr, ss = u[idx], r[idx], s[idx]
        inter = kn.pack_interactions(uu, np.zeros(n, dtype=np.int64), rr, cuda_dev)
        order, seg = kn.user_segments(uu)
        assert (order is not None) == shuffle
        out = kn.rank_metrics(inter, torch.tensor(ss, device=cuda_dev), torch.tensor(seg, device=cuda_dev),
                              None if order is None else torch.tensor(order, device=cuda_dev)).cpu().numpy()
        nd, hr = evalm.rank_metrics(uu, rr, ss, kind="stable")
        assert out[2] == len(lens) and abs(out[0] / len(lens) - nd) < 1e-12 and abs(out[1] / len(lens) - hr) < 1e-12


def test_routing_and_merge_vs_reference_golden(cuda_dev):
    """retrain_gid and merged tables bit-exact (Appendix E)."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    z = load_gold("toy_sisa.npz")
    K, U = int(z["K"]), 1508
    groups = [z[f"group{s}"] for s in range(K)]
    owner = np.full(U, -1, dtype=np.int32)
    for s, g in enumerate(groups):
        owner[g] = s
    t_owner = torch.tensor(owner, device=cuda_dev)
    flags = kn.route_deletions(t_owner, torch.tensor(z["del_user"].astype(np.int32), device=cuda_dev), K)
    assert np.flatnonzero(flags.cpu().numpy()).tolist() == z["retrain_gid"].tolist()
    assert set(np.flatnonzero(flags.cpu().numpy())) == osisa.route_deletions(groups, z["del_user"])
    # merge: learn (zero + owner rows) then unlearn (clone + retrained owners)
    rng = np.random.default_rng(2)
    Ps = [rng.standard_normal((U, 16), dtype=np.float32) for _ in range(K)]
    tP = [torch.tensor(p, device=cuda_dev) for p in Ps]
    merged = torch.full((U, 16), 7.0, device=cuda_dev)
    kn.merge_user_rows(tP, t_owner, merged, zero_unowned=True)
    assert np.array_equal(merged.cpu().numpy(), osisa.merge_learn(Ps, groups))
    P2 = [rng.standard_normal((U, 16), dtype=np.float32) for _ in range(K)]
    before = merged.cpu().numpy().copy()
    kn.merge_user_rows([torch.tensor(p, device=cuda_dev) for p in P2], t_owner, merged, retrain=flags)
    assert np.array_equal(merged.cpu().numpy(), osisa.merge_unlearn(before, P2, groups, z["retrain_gid"].tolist()))


@pytest.mark.parametrize("n,d,k", [(6040, 16, 5), (1000, 8, 3), (5000, 64, 32), (3001, 32, 8), (2048, 128, 16),
                                   (4096, 64, 128), (700, 64, 200)])
def test_cost_matrix_tcgen05_vs_float64(cuda_dev, n, d, k):
    """tcgen05 3xTF32 cost == float64 sum (x-c)^2 within 1e-5 relative (Appendix E); also vs the SIMT kernel."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(n + d + k)
    X = rng.standard_normal((n, d), dtype=np.float32)
    Cc = X[rng.choice(n, k, replace=False)] + 0.1 * rng.standard_normal((k, d), dtype=np.float32)
    tX, tC = torch.tensor(X, device=cuda_dev), torch.tensor(Cc, device=cuda_dev)
    M, inertia = kn.cost_matrix(tX, tC, want_inertia=True)
    Ms, inertia_s = kn.cost_matrix(tX, tC, want_inertia=True, simt=True)
    ref = oot.cost_matrix(X, Cc)
    M, Ms = M.cpu().numpy(), Ms.cpu().numpy()
    assert np.isinf(M[:, k:]).all() and np.isinf(Ms[:, k:]).all()
    scale = ref.max()
    assert np.abs(Ms[:, :k] - ref).max() / scale < 1e-5
    assert np.abs(M[:, :k] - ref).max() / scale < 1e-5
    ref_in = ref.min(axis=1).sum()
    assert abs(inertia.item() - ref_in) / ref_in < 1e-5 and abs(inertia_s.item() - ref_in) / ref_in < 1e-5


@pytest.mark.parametrize("n,k", [(6040, 5), (999, 3), (20000, 32), (4000, 128), (3000, 200)])
def test_sinkhorn_plan_vs_float64_oracle(cuda_dev, n, k):
    """Persistent Sinkhorn: n*P within 1e-4 of the float64 oracle at the same eps schedule (Appendix E);
    the split-phase kernels (multi-GPU path) give the same potentials; labels == argmax of the oracle plan."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(n + k)
    d = 16
    X = rng.standard_normal((n, d), dtype=np.float32)
    Cc = X[rng.choice(n, k, replace=False)]
    M = kn.cost_matrix(torch.tensor(X, device=cuda_dev), torch.tensor(Cc, device=cuda_dev))
    Mh = M.cpu().numpy()[:, :k].astype(np.float64)
    mean = float(Mh.mean())
    sched = [(mean * 0.5, 20), (mean * 0.1, 40), (mean * 0.03, 60)]
    g = kn.sinkhorn(M, k, sched)
    P_o, _, g_o, _ = oot.sinkhorn_log(Mh, sched)
    assert np.abs(g.cpu().numpy() - g_o).max() < 1e-3 * mean
    plan = kn.sinkhorn_plan(M, k, g, sched[-1][0]).cpu().numpy()
    assert np.abs(n * plan - n * P_o).max() < 1e-4
    np.testing.assert_allclose(plan.sum(1), 1.0 / n, rtol=1e-5)
    # split phase (what the multi-GPU driver runs, with an all-reduce between the two kernels)
    g2 = torch.zeros(k, dtype=torch.float32, device=cuda_dev)
    colsum = torch.zeros(M.shape[1], dtype=torch.float64, device=cuda_dev)
    for eps, iters in sched:
        for _ in range(iters):
            kn.sinkhorn_colsum(M, k, g2, eps, n, colsum)
            kn.sinkhorn_update_g(g2, colsum, k, eps)
    assert (g2 - g).abs().max().item() < 1e-4 * mean
    # assignment: bit-exact given the same plan (oracle plan in fp64 through our extraction kernel) ...
    lab_o = oot.assign(P_o)
    lab = kn.assign_plan(torch.tensor(P_o, device=cuda_dev)).cpu().numpy()
    assert np.array_equal(lab, lab_o)
    # ... and the fused potentials->label kernel agrees with argmax of our own plan
    label, sums, cnt = kn.assign_centroids(M, k, g, torch.tensor(X, device=cuda_dev))
    label = label.cpu().numpy()
    assert np.array_equal(label, np.argmax(g.cpu().numpy()[None, :] - M.cpu().numpy()[:, :k], axis=1))
    assert (label == lab_o).mean() > 0.999
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(label, minlength=k))
    ref_sum = np.stack([X[label == j].astype(np.float64).sum(0) for j in range(k)])
    assert np.abs(sums.cpu().numpy() - ref_sum).max() < 1e-3


@pytest.mark.parametrize("form", ["cluster", "persistent", "split"])
def test_sinkhorn_dead_column_recovers_exactly_like_the_log_domain_oracle(cuda_dev, form):
    """Potentials that are stale by hundreds of eps (a warm start after the centroids moved, profiles/r1_notes.md):
    every entry of one column is > 700 binades below its row maximum, so a linear-domain fp32 column sum is exactly
    zero and log(0) sends the potential to infinity.  The kernels keep the column sums in the log domain: all three
    forms follow the float64 log-sum-exp oracle from the very first iteration."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(11)
    n, k = {"cluster": (6040, 5), "persistent": (20000, 32), "split": (20000, 32)}[form]
    X = rng.standard_normal((n, 16), dtype=np.float32)
    Cc = X[rng.choice(n, k, replace=False)]
    M = kn.cost_matrix(torch.tensor(X, device=cuda_dev), torch.tensor(Cc, device=cuda_dev))
    Mh = M.cpu().numpy()[:, :k].astype(np.float64)
    mean = float(Mh.min(1).mean())
    eps = 0.02 * mean
    g0 = np.zeros(k, dtype=np.float32)
    g0[1] = -12.0 * mean                                   # (g_1 - M_i1)/eps is ~600 nats (870 binades) below every
    g0[k - 1] = -3.0 * mean                                # row's maximum: far beyond an fp32 linear-domain sum
    sched = [(eps, 1)]
    _, _, g_o1, _ = oot.sinkhorn_log(Mh, sched, g0=g0.astype(np.float64))
    assert np.isfinite(g_o1).all() and g_o1[1] - g0[1] > 10.0 * mean
    sched = [(eps, 40)]
    P_o, _, g_o, _ = oot.sinkhorn_log(Mh, sched, g0=g0.astype(np.float64))

    def run(iters):
        g = torch.tensor(g0, device=cuda_dev)
        if form == "split":
            colsum = torch.zeros(M.shape[1], dtype=torch.float64, device=cuda_dev)
            for _ in range(iters):
                kn.sinkhorn_colsum(M, k, g, eps, n, colsum)
                kn.sinkhorn_update_g(g, colsum, k, eps)
            return g
        return kn.sinkhorn(M, k, [(eps, iters)], g=g)

    g1 = run(1).cpu().numpy()
    assert np.isfinite(g1).all()
    assert np.abs(g1 - g_o1).max() < 2e-3 * mean, (g1, g_o1)           # the dead columns are lifted by the exact amount
    g = run(40)
    plan = kn.sinkhorn_plan(M, k, g, eps).cpu().numpy()
    assert np.abs(n * plan - n * P_o).max() < 2e-4
    lab = np.argmax(g.cpu().numpy()[None, :] - M.cpu().numpy()[:, :k], axis=1)
    assert np.bincount(lab, minlength=k).min() > 0.5 * n / k


def test_sinkhorn_nan_cost_column_propagates(cuda_dev):
    """A NaN centroid (an empty group upstream) must not be absorbed by a floor: the potentials come back non-finite
    and ot_cluster_device raises instead of caching a degenerate grouping."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(3)
    X = rng.standard_normal((4000, 16), dtype=np.float32)
    Cc = X[:5].copy()
    Cc[2] = np.nan
    M = kn.cost_matrix(torch.tensor(X, device=cuda_dev), torch.tensor(Cc, device=cuda_dev))
    g = kn.sinkhorn(M, 5, [(5.0, 5)])
    assert not np.isfinite(g.cpu().numpy()).all()


@pytest.mark.parametrize("n,k,d", [(6040, 5, 16), (1001, 7, 8), (30000, 8, 16), (50000, 32, 16), (9000, 128, 32)])
def test_balanced_rounding_vs_oracle_and_exact_emd(cuda_dev, n, k, d):
    """ure_balance_labels on the argmax of our own Sinkhorn potentials: bit-equal to the oracle's successive-shortest-
    path rounding on the same cost matrix and labels (single-CTA and multi-CTA launches, k | n and not), every group
    floor(n/k)..ceil(n/k) users; at ml1m size the result is the argmax of the exact LP plan (the reference's ot.emd,
    utils.py:644-647)."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(n + k)
    X = rng.standard_normal((n, d), dtype=np.float32)
    Cc = X[rng.choice(n, k, replace=False)]
    Xd = torch.tensor(X, device=cuda_dev)
    M = kn.cost_matrix(Xd, torch.tensor(Cc, device=cuda_dev))
    Mh = M.cpu().numpy()[:, :k]
    scale = float(Mh.min(1).mean())
    g = kn.sinkhorn(M, k, [(scale, 10), (0.3 * scale, 20), (0.1 * scale, 30), (0.03 * scale, 60)])
    label, _, cnt = kn.assign_centroids(M, k, g, None)
    lab0 = label.cpu().numpy().copy()
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(lab0, minlength=k))
    status = kn.balance_labels(M, k, label, cnt)
    got, sizes, status = label.cpu().numpy(), cnt.cpu().numpy(), status.cpu().numpy()
    lo, hi = n // k, -(-n // k)
    assert sizes.min() >= lo and sizes.max() <= hi and sizes.sum() == n, sizes
    assert np.array_equal(sizes, np.bincount(got, minlength=k)) and status[1] == 0 and status[2] == 0
    ref, n_aug = oot.balance_labels(Mh, lab0)
    assert status[0] == n_aug and n_aug > 0
    assert np.array_equal(got, ref)
    sums, cnt2 = kn.centroid_sums(Xd, label, k)
    assert np.array_equal(cnt2.cpu().numpy(), sizes)
    ref_sum = np.stack([X[got == j].astype(np.float64).sum(0) for j in range(k)])
    assert np.abs(sums.cpu().numpy() - ref_sum).max() < 1e-3
    if n == 6040:
        emd = oot.assign(oot.emd_lp(np.ones(n) / n, np.ones(k) / k, Mh))
        assert (got == emd).mean() >= 0.9995, (got != emd).sum()
        rows = np.arange(n)
        assert abs(float(Mh[rows, got].sum()) - float(Mh[rows, emd].sum())) <= 1e-5 * float(Mh[rows, emd].sum())


def test_assign_plan_first_max_wins_on_emd_plan(cuda_dev):
    """The reference's plan (exact EMD vertex) through the extraction kernel: bit-exact labels, ties -> first."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(5)
    n, k = 600, 4
    M = rng.random((n, k))
    G = oot.emd_lp(np.ones(n) / n, np.ones(k) / k, M)
    G[:7] = 0.0                                             # all-tied rows -> label 0 like np.argmax
    lab = kn.assign_plan(torch.tensor(G, device=cuda_dev)).cpu().numpy()
    assert np.array_equal(lab, np.argmax(G, axis=1))
    lab32 = kn.assign_plan(torch.tensor(G.astype(np.float32), device=cuda_dev)).cpu().numpy()
    assert np.array_equal(lab32, np.argmax(G.astype(np.float32), axis=1))


def test_errors_are_loud(cuda_dev):
    torch = _torch()
    from ultrare_b200 import kernels as kn
    with pytest.raises(RuntimeError):
        kn.cost_matrix(torch.zeros((10, 12), device=cuda_dev), torch.zeros((2, 12), device=cuda_dev))   # d=12
    with pytest.raises(RuntimeError):
        kn.ensemble_score([torch.zeros((4, 16))], [torch.zeros((4, 16))], torch.zeros((1, 4), dtype=torch.int32))


@pytest.mark.parametrize("d,K", [(16, 1), (64, 3), (128, 2)])
def test_mf_train_lazy_equals_dense_reference_arithmetic(cuda_dev, d, K):
    """Lazy mode (closed-form catch-up of untouched rows, M^n table) == the reference's dense optimiser:
    tables far larger than a batch, so most rows are untouched most steps; compared with the dense NumPy
    oracle after ure_mf_flush (weights 1e-4 abs, losses 1e-5 rel), ragged shards included."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(100 + d)
    U, I, batch, epochs = 3000, 2500, 512, 3
    shards, host = [], []
    for s in range(K):
        n = 6000 + 1700 * s
        u, i = rng.integers(0, U, n), rng.integers(0, I, n)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
        Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
        perms = [omf.feistel_perm(n, omf.perm_key(9, s, ep)) for ep in range(epochs)]
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                                    torch.tensor(Q0, device=cuda_dev), epochs, s, 9))
        host.append((u, i, r, P0, Q0, perms))
    sb = kn.ShardBatch(shards, d, batch, lazy=True)
    sb.train()
    sb.flush()
    torch.cuda.synchronize()
    losses = sb.train_losses()
    for s in range(K):
        u, i, r, P0, Q0, perms = host[s]
        P, Q, bP, bQ, ls = omf.mf_train(P0, Q0, u, i, r, perms, batch, epochs)
        np.testing.assert_allclose(losses[s], ls, rtol=1e-5)
        assert np.abs(shards[s].P.cpu().numpy() - P).max() < 1e-4
        assert np.abs(shards[s].Q.cpu().numpy() - Q).max() < 1e-4
        assert np.abs(shards[s].bufP.cpu().numpy() - bP).max() < 1e-3
        assert float(shards[s].gP.abs().max()) == 0.0 and float(shards[s].gQ.abs().max()) == 0.0
        untouched = np.setdiff1d(np.arange(U), u)
        assert len(untouched) > 0                      # rows that only ever decayed are exact too
        assert np.abs(shards[s].P.cpu().numpy()[untouched] - P[untouched]).max() < 1e-5


@pytest.mark.parametrize("d,K,explicit,list_bytes", [(128, 8, False, None), (64, 3, False, 1), (16, 2, True, None),
                                                       (128, 1, True, 1), (8, 2, False, None)])
def test_mf_train_runs_equals_dense_reference_arithmetic(cuda_dev, monkeypatch, d, K, explicit, list_bytes):
    """RUNS schedule (csrc/mf_train_runs.cu: owner-computes for tables in HBM -- sorted step lists, whole runs of a
    row per warp, two row versions + a tag, closed-form catch-up of untouched rows) == the reference's dense
    optimiser.  c4small-like proportions (tables far larger than a batch, d = 128, K = 8 shards, 2 epochs; SURVEY
    config C4 scaled down so that the dense NumPy oracle finishes), ragged shards, skewed rows (long runs),
    explicit visiting orders, one-epoch list windows; after ure_mf_runs_flush weights 1e-4 abs, losses 1e-5 rel."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    if list_bytes is not None:
        monkeypatch.setattr(kn, "RUNS_LIST_BYTES", list_bytes)
    rng = np.random.default_rng(300 + d + K)
    U, I, batch, epochs = 2500, 5000, 700, 2
    shards, host = [], []
    for s in range(K):
        n = 9000 + 2300 * s
        u = (rng.random(n) ** 1.5 * U).astype(np.int64)                # heavy-tailed rows, as synth.device_interactions
        i = (rng.random(n) ** 2.0 * I).astype(np.int64)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
        Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
        if explicit:
            perms = [rng.permutation(n) for _ in range(epochs)]
            perm_t = torch.tensor(np.stack(perms).astype(np.int32), device=cuda_dev)
        else:
            perms = [omf.feistel_perm(n, omf.perm_key(9, s, ep)) for ep in range(epochs)]
            perm_t = None
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                                    torch.tensor(Q0, device=cuda_dev), epochs, s, 9, perm=perm_t))
        host.append((u, i, r, P0, Q0, perms))
    sb = kn.ShardBatch(shards, d, batch, mode="runs")
    assert sb.mode == "runs" and sb.hp.runs_rows == (1 if list_bytes else epochs)
    sb.train()
    sb.flush()
    torch.cuda.synchronize()
    losses = sb.train_losses()
    for s in range(K):
        u, i, r, P0, Q0, perms = host[s]
        P, Q, bP, bQ, ls = omf.mf_train(P0, Q0, u, i, r, perms, batch, epochs)
        np.testing.assert_allclose(losses[s], ls, rtol=1e-5)
        assert np.abs(shards[s].P.cpu().numpy() - P).max() < 1e-4
        assert np.abs(shards[s].Q.cpu().numpy() - Q).max() < 1e-4
        assert np.abs(shards[s].bufP.cpu().numpy() - bP).max() < 1e-3
        assert np.abs(shards[s].bufQ.cpu().numpy() - bQ).max() < 1e-3
        untouched = np.setdiff1d(np.arange(I), i)
        assert len(untouched) > 0                      # rows that only ever decayed are exact too
        assert np.abs(shards[s].Q.cpu().numpy()[untouched] - Q[untouched]).max() < 1e-5


def test_mf_train_runs_epoch_by_epoch_equals_single_launch(cuda_dev):
    """RUNS: training in two launches (the row slots and version tags persist between them) == one launch."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(17)
    U, I, n, d, batch, epochs = 900, 1500, 7000, 32, 400, 3
    u, i = (rng.random(n) ** 1.5 * U).astype(np.int64), (rng.random(n) ** 2.0 * I).astype(np.int64)
    r = rng.integers(1, 6, n).astype(np.float32) / 5
    P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
    Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
    outs = []
    for cuts in ([None], [5, 18, 19, None]):
        st = kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                           torch.tensor(Q0, device=cuda_dev), epochs, 1, 9)
        sb = kn.ShardBatch([st], d, batch, mode="runs")
        for c in cuts:
            sb.train(c)
            sb.flush()                                 # exporting the tables in between must not disturb the slots
        outs.append((st.P.cpu().numpy(), st.Q.cpu().numpy(), sb.train_losses()[0]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-6)


def test_mf_owner_prepare_sorts_and_inverts(cuda_dev):
    """ure_mf_owner_prepare: the user-sorted / item-sorted record copies are permutations of the shard's records
    grouped by row (pad = original index), the offsets are the row histograms' prefix sums, perm_inv inverts perm."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(5)
    U, I, n, epochs = 70, 90, 4000, 2
    u, i = rng.integers(0, U, n), rng.integers(0, I - 7, n)        # items I-7.. never occur: empty trailing rows
    r = rng.integers(1, 6, n).astype(np.float32) / 5
    perm = np.stack([rng.permutation(n) for _ in range(epochs)]).astype(np.int32)
    st = kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.zeros((U, 16), device=cuda_dev),
                       torch.zeros((I, 16), device=cuda_dev), epochs, perm=torch.tensor(perm, device=cuda_dev))
    sb = kn.ShardBatch([st], 16, 512, mode="owner")
    torch.cuda.synchronize()
    for rec, off, key, rows in ((st.inter_u, st.off_u, u, U), (st.inter_i, st.off_i, i, I)):
        rec, off = rec.cpu().numpy(), off.cpu().numpy()
        assert np.array_equal(off[:rows + 1], np.concatenate([[0], np.cumsum(np.bincount(key, minlength=rows))]))
        assert np.array_equal(np.sort(rec[:, 3]), np.arange(n))
        j = rec[:, 3]
        assert np.array_equal(rec[:, 0], u[j]) and np.array_equal(rec[:, 1], i[j])
        assert np.array_equal(rec[:, 2].view(np.float32), r[j])
        col = rec[:, 0] if key is u else rec[:, 1]
        assert np.all(np.diff(col) >= 0)
        assert np.array_equal(j, np.argsort(key, kind="stable"))       # stable: record order inside every row
    inv = st.perm_inv.cpu().numpy()
    for e in range(epochs):
        assert np.array_equal(inv[e][perm[e]], np.arange(n))
    assert sb.owner_plan["smem_need"] <= sb.owner_plan["smem_avail"]


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("U,I,n", [(200, 250, 20_000), (5000, 3000, 150_000), (70_000, 300, 90_000), (40, 70_000, 60_001)])
def test_mf_owner_setup_one_launch_equals_eight(cuda_dev, monkeypatch, fused, U, I, n):
    """URE_SETUP_FUSED: the whole set-up as ONE cooperative launch (owner_setup_kernel: the same bodies as virtual
    blocks, grid barriers where the launch boundaries were) against the eight separate launches -- one, two and three
    radix passes, row counters in shared memory and (tables too large) in global memory, two shards of which one is
    already in user order.  Both give the stable argsort and the row histograms' prefix sums."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    monkeypatch.setenv("URE_SETUP_FUSED", fused)
    rng = np.random.default_rng(U + n)
    shards, host = [], []
    for kind in ("random", "sorted"):
        u = rng.integers(0, U, n) if kind == "random" else np.sort(rng.integers(0, U, n))
        i = rng.integers(0, I, n)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.zeros((U, 16), device=cuda_dev),
                                    torch.zeros((I, 16), device=cuda_dev), 1))
        host.append((u, i))
    try:
        kn.ShardBatch(shards, 16, 4096, mode="owner")
    except RuntimeError as e:              # the set-up has run; the training plan of the widest tables does not fit
        assert "does not fit" in str(e) and max(U, I) > 50_000
    torch.cuda.synchronize()
    for st, (u, i) in zip(shards, host):
        for rec, off, key, rows in ((st.inter_u, st.off_u, u, U), (st.inter_i, st.off_i, i, I)):
            rec, off = rec.cpu().numpy(), off.cpu().numpy()
            j = np.argsort(key, kind="stable")
            assert np.array_equal(rec[:, 3], j) and np.array_equal(rec[:, 0], u[j]) and np.array_equal(rec[:, 1], i[j])
            assert np.array_equal(off[:rows + 1], np.concatenate([[0], np.cumsum(np.bincount(key, minlength=rows))]))


def test_mf_owner_prepare_records_already_in_user_order(cuda_dev):
    """Rating files are written user by user: a shard whose records arrive in user-row order skips the user-side
    radix passes (the counting pass notices and keeps the copy it wrote while reading).  Three shards in one batch --
    in order, in order except for ONE late inversion, random -- give the same sorted copies as a stable argsort, and
    the batched runtime's arena path agrees."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(9)
    U, I, d = 300, 500, 16
    shards, host = [], []
    for kind, n in (("sorted", 70_001), ("one inversion", 50_000), ("random", 30_000)):
        u = np.sort(rng.integers(0, U, n)) if kind != "random" else rng.integers(0, U, n)
        if kind == "one inversion":
            u[-1] = 0
        i = rng.integers(0, I, n)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.zeros((U, d), device=cuda_dev),
                                    torch.zeros((I, d), device=cuda_dev), 1))
        host.append((u, i))
    kn.ShardBatch(shards, d, 4096, mode="owner")
    torch.cuda.synchronize()
    for st, (u, i) in zip(shards, host):
        for rec, key in ((st.inter_u, u), (st.inter_i, i)):
            rec = rec.cpu().numpy()
            j = np.argsort(key, kind="stable")
            assert np.array_equal(rec[:, 3], j) and np.array_equal(rec[:, 0], u[j]) and np.array_equal(rec[:, 1], i[j])
    sb = kn.ArenaShardBatch([st.inter for st in shards], [U] * 3, I, d, 4096, 1, [1, 2, 3])
    torch.cuda.synchronize()
    for v, (u, i) in zip(sb.shards, host):
        pass                                              # (the views expose tables only; the arena's sorted copies:)
    lay, n_tot = sb._lay, sum(len(h[0]) for h in host)
    rec = sb.arena[int(lay.rec):int(lay.rec) + 2 * n_tot * 16].view(torch.int32).view(2, n_tot, 4).cpu().numpy()
    o = 0
    for u, i in host:
        n = len(u)
        assert np.array_equal(rec[0, o:o + n, 3], np.argsort(u, kind="stable"))
        assert np.array_equal(rec[1, o:o + n, 3], np.argsort(i, kind="stable"))
        o += n


@pytest.mark.parametrize("cache,force", [(True, None), (False, None), (False, (2, 64)), (False, (0, 16)), (False, (2, 4096))])
def test_mf_owner_skewed_rows_many_shards_vs_oracle(cuda_dev, cache, force, monkeypatch):
    """Owner schedule on ragged shards with heavy rows (one user / one item holding a large share of a shard, rows
    split over many chunks and CTAs with no rows at all), more steps per epoch than a chunk window, empty shard.
    force = (owner_flags, list_cap): the large-problem configurations -- batch lists longer than the staged
    capacity are read from the schedule table, the pre-pass runs without its record-index cache."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(77)
    d, batch, epochs, K = 16, 257, 2, 7
    shards, host = [], []
    for s in range(K):
        U, I = 40 + 13 * s, 55 + 9 * s
        n = 0 if s == 3 else 3000 + 2500 * s
        u, i = rng.integers(0, U, n), rng.integers(0, I, n)
        if n:
            u[rng.random(n) < 0.4] = 3                  # a user with 40 % of the shard
            i[rng.random(n) < 0.3] = 5
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
        Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
        perms = [omf.feistel_perm(n, omf.perm_key(9, s, ep)) for ep in range(epochs)] if n else []
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                                    torch.tensor(Q0, device=cuda_dev), epochs, s, 9))
        host.append((u, i, r, P0, Q0, perms))
    monkeypatch.setattr(kn, "OWNER_FORCE", force)
    sb = kn.ShardBatch(shards, d, batch, mode="owner", owner_cache=cache)
    assert sb.owner_plan["cached"] == cache
    if force is not None:
        assert sb.owner_plan["flags"] == force[0] and sb.owner_plan["list_cap"] == min(force[1], sb.hp.owner_cap_slots)
    sb.train()
    torch.cuda.synchronize()
    losses = sb.train_losses()
    for s in range(K):
        u, i, r, P0, Q0, perms = host[s]
        if len(u) == 0:
            assert np.array_equal(shards[s].P.cpu().numpy(), P0)
            continue
        P, Q, bP, bQ, ls = omf.mf_train(P0, Q0, u, i, r, perms, batch, epochs)
        np.testing.assert_allclose(losses[s], ls, rtol=1e-5)
        assert np.abs(shards[s].P.cpu().numpy() - P).max() < 1e-4
        assert np.abs(shards[s].Q.cpu().numpy() - Q).max() < 1e-4
        assert np.abs(shards[s].bufQ.cpu().numpy() - bQ).max() < 1e-3
        assert float(shards[s].gP.abs().max()) == 0.0 and float(shards[s].gQ.abs().max()) == 0.0


@pytest.mark.parametrize("mode,sched_bytes,explicit", [("owner", None, False), ("owner", 1, False), ("dense", None, False),
                                                       ("owner", None, True)])
def test_arena_batch_runtime_equals_per_shard_states(cuda_dev, monkeypatch, mode, sched_bytes, explicit):
    """The native batch runtime (ure_mf_batch_layout / _setup / _plan, kernels.ArenaShardBatch) against the per-shard
    ShardState path on the same records, weights and visiting orders: identical descriptors in effect -- the trained
    tables, momentum buffers and per-epoch losses are equal bit for bit (owner) / to rounding (dense: atomics), and
    equal to the oracle.  Ragged shards, an empty shard, schedule windows of 2 epochs, explicit visiting orders."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    if sched_bytes is not None:
        monkeypatch.setattr(kn, "OWNER_SCHED_BYTES", sched_bytes)
    rng = np.random.default_rng(77)
    d, batch, epochs, I = 16, 900, 4, 150
    sizes = [(110, 5200), (60, 0), (140, 7900), (95, 3100)]
    recs, perms_h, host = [], [], []
    for s, (U, n) in enumerate(sizes):
        u, i = rng.integers(0, U, n), rng.integers(0, I, n)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        recs.append(kn.pack_interactions(u, i, r, cuda_dev))
        if explicit:
            perms_h.append(np.stack([rng.permutation(n) for _ in range(epochs)]).astype(np.int32).reshape(epochs, n))
        else:
            perms_h.append([omf.feistel_perm(n, omf.perm_key(9, s + 1, ep)) for ep in range(epochs)] if n else [])
        host.append((u, i, r))
    perms_d = [torch.tensor(p, device=cuda_dev) for p in perms_h] if explicit else None
    gen = torch.Generator(device=cuda_dev).manual_seed(5)
    ab = kn.ArenaShardBatch(recs, [U for U, _ in sizes], I, d, batch, epochs, [s + 1 for s in range(len(sizes))], 9,
                            perms_d, generator=gen, std=0.3, mode=mode)
    assert ab.mode == mode
    P0 = [st.P.clone() for st in ab.shards]
    Q0 = [st.Q.clone() for st in ab.shards]
    ab.train()
    la = ab.train_losses()
    states = [kn.ShardState(recs[s], P0[s].clone(), Q0[s].clone(), epochs, s + 1, 9,
                            perm=None if perms_d is None else perms_d[s]) for s in range(len(sizes))]
    sb = kn.ShardBatch(states, d, batch, mode=mode)
    sb.train()
    lb = sb.train_losses()
    for s, (U, n) in enumerate(sizes):
        a, b = ab.shards[s], states[s]
        if mode == "owner":
            assert torch.equal(a.P, b.P) and torch.equal(a.Q, b.Q) and torch.equal(a.bufP, b.bufP)
            assert np.array_equal(la[s], lb[s])
        else:
            assert (a.P - b.P).abs().max().item() < 1e-5 and (a.Q - b.Q).abs().max().item() < 1e-5
        if n == 0:
            assert torch.equal(a.P, P0[s])
            continue
        u, i, r = host[s]
        P, Q, bP, bQ, ls = omf.mf_train(P0[s].cpu().numpy(), Q0[s].cpu().numpy(), u, i, r, list(perms_h[s]), batch, epochs)
        np.testing.assert_allclose(la[s], ls, rtol=1e-5)
        assert np.abs(a.P.cpu().numpy() - P).max() < 1e-4 and np.abs(a.Q.cpu().numpy() - Q).max() < 1e-4
        assert float(a.gP.abs().max()) == 0.0 and float(a.gQ.abs().max()) == 0.0


def test_pack_interactions_on_device_equals_host_casts(cuda_dev):
    """ure_pack_interactions_f64 == RatingData's host casts (read.py:111-113,124): int(uid), int(iid),
    float32(float64 rating), bit for bit, with and without the compact-row mapping; empty input."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(3)
    n, U = 200_003, 5000
    raw = np.vstack([rng.integers(0, U, n).astype(np.float64), rng.integers(0, 3000, n).astype(np.float64),
                     rng.integers(1, 11, n) / 2.0 / 5.0])
    row_of = rng.permutation(U).astype(np.int32)
    got = kn.upload_interactions(raw, cuda_dev).cpu().numpy()
    assert np.array_equal(got[:, 0], raw[0].astype(int)) and np.array_equal(got[:, 1], raw[1].astype(int))
    assert np.array_equal(got[:, 2].view(np.float32), raw[2].astype(np.float32)) and not got[:, 3].any()
    got = kn.upload_interactions(raw, cuda_dev, torch.tensor(row_of, device=cuda_dev)).cpu().numpy()
    assert np.array_equal(got[:, 0], row_of[raw[0].astype(int)])
    assert np.array_equal(got, kn.pack_interactions(row_of[raw[0].astype(int)], raw[1], raw[2], cuda_dev).cpu().numpy())
    assert kn.upload_interactions(np.zeros((3, 0)), cuda_dev).shape == (0, 4)


def test_uploads_through_the_staging_ring_equal_one_shot_uploads(cuda_dev, monkeypatch):
    """Tables / rating arrays beyond UPLOAD_ONE_SHOT_BYTES travel through a two-segment page-locked ring (a C4-sized
    table cannot be staged whole).  With the limits lowered: ragged sizes (not a multiple of a segment, of 2 MB, of 16
    bytes), several arrays with an empty one and a page-locked one among them, deferred and direct -- bit-equal to
    the one-shot path."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(5)
    tab = rng.standard_normal(3 * (1 << 20) + 12345)
    small = rng.integers(0, 1 << 30, 700_001).astype(np.int32)
    ns, U = [300_007, 0, 150_001, 90_000], 4000
    raws = [np.vstack([rng.integers(0, U, n).astype(np.float64), rng.integers(0, 3000, n).astype(np.float64),
                       rng.integers(1, 11, n) / 2.0 / 5.0]) for n in ns]
    raws[3] = kn.pinned_copy(raws[3])
    row_of = torch.tensor(rng.permutation(U).astype(np.int32), device=cuda_dev)
    ref_tab, ref_small = kn.upload_table(tab, cuda_dev), kn.upload_table(small, cuda_dev)
    ref = kn.upload_interactions_many(raws, cuda_dev, row_of)
    monkeypatch.setattr(kn, "UPLOAD_ONE_SHOT_BYTES", 1 << 20)
    monkeypatch.setattr(kn, "UPLOAD_RING_BYTES", (1 << 22) + (1 << 20))       # 5 MB: segments of 2.5 staging tasks
    assert torch.equal(kn.upload_table(tab, cuda_dev), ref_tab) and np.array_equal(ref_tab.cpu().numpy(), tab)
    assert torch.equal(kn.upload_table(small, cuda_dev), ref_small)
    got = kn.upload_interactions_many(raws, cuda_dev, row_of)
    outs, finish = kn.upload_interactions_many(raws, cuda_dev, row_of, defer=True)
    finish()
    for a, b, c in zip(ref, got, outs):
        assert torch.equal(a, b) and torch.equal(a, c)
    assert got[1].shape == (0, 4)


def test_mf_owner_schedule_windows_vs_oracle(cuda_dev, monkeypatch):
    """Owner schedule with schedule tables that hold only 2 epochs per shard: the training is split into windows,
    each with its own ure_mf_owner_schedule pass; ragged shards (different steps per epoch), 5 epochs."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    monkeypatch.setattr(kn, "OWNER_SCHED_BYTES", 1)
    rng = np.random.default_rng(21)
    d, batch, epochs, K = 16, 700, 5, 3
    shards, host = [], []
    for s in range(K):
        U, I, n = 90 + 20 * s, 120, 4000 + 1500 * s
        u, i = rng.integers(0, U, n), rng.integers(0, I, n)
        r = rng.integers(1, 6, n).astype(np.float32) / 5
        P0 = rng.standard_normal((U, d), dtype=np.float32) * 0.3
        Q0 = rng.standard_normal((I, d), dtype=np.float32) * 0.3
        perms = [omf.feistel_perm(n, omf.perm_key(5, s, ep)) for ep in range(epochs)]
        shards.append(kn.ShardState(kn.pack_interactions(u, i, r, cuda_dev), torch.tensor(P0, device=cuda_dev),
                                    torch.tensor(Q0, device=cuda_dev), epochs, s, 5))
        host.append((u, i, r, P0, Q0, perms))
    sb = kn.ShardBatch(shards, d, batch, mode="owner")
    assert sb.owner_plan["schedule_rows"] == 2
    sb.train()
    torch.cuda.synchronize()
    losses = sb.train_losses()
    for s in range(K):
        u, i, r, P0, Q0, perms = host[s]
        P, Q, _, _, ls = omf.mf_train(P0, Q0, u, i, r, perms, batch, epochs)
        np.testing.assert_allclose(losses[s], ls, rtol=1e-5)
        assert np.abs(shards[s].P.cpu().numpy() - P).max() < 1e-4
        assert np.abs(shards[s].Q.cpu().numpy() - Q).max() < 1e-4


@pytest.mark.parametrize("n,d,k", [(5003, 64, 8), (7001, 128, 16), (4099, 32, 32), (3000, 64, 40), (2000, 8, 5)])
def test_assign_centroids_register_and_shared_paths(cuda_dev, n, d, k):
    """ure_assign_centroids: label = argmax_j (g_j - M_ij) (first max wins), per-label counts and fp64 sums of X,
    on the register path (k <= 32, k*ceil(d/32) <= 64) and the shared-memory path."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(n + d + k)
    X = rng.standard_normal((n, d), dtype=np.float32)
    Cc = X[rng.choice(n, k, replace=False)]
    Xd = torch.tensor(X, device=cuda_dev)
    M = kn.cost_matrix(Xd, torch.tensor(Cc, device=cuda_dev))
    g = torch.tensor(rng.standard_normal(k).astype(np.float32), device=cuda_dev)
    label, sums, cnt = kn.assign_centroids(M, k, g, Xd)
    label = label.cpu().numpy()
    ref = np.argmax(g.cpu().numpy()[None, :] - M.cpu().numpy()[:, :k], axis=1)
    assert np.array_equal(label, ref)
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(ref, minlength=k))
    ref_sum = np.stack([X[ref == j].astype(np.float64).sum(0) for j in range(k)])
    assert np.abs(sums.cpu().numpy() - ref_sum).max() < 2e-3


def test_mf_owner_training_is_bit_reproducible(toy, cuda_dev):
    """Owner schedule: no atomics on the gradient path and a stable sort in the set-up, so two trainings of the same
    shards give bit-identical tables (the dense schedule's L2 atomics do not)."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    outs = []
    for rep in range(2):
        shards, _ = _shard(2, toy, cuda_dev, 2, 3000, False, [3, 4])
        sb = kn.ShardBatch(shards, toy["k"], 3000, mode="owner")
        sb.train()
        torch.cuda.synchronize()
        outs.append([t.cpu().numpy().copy() for sh in shards for t in (sh.P, sh.Q, sh.bufP, sh.bufQ)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", MODES)
def test_mf_full_size_ml1m_k5_vs_oracle(cuda_dev, mode):
    """BASELINE config 2 at its full size (ml1m shape: 6040 users, 3416 items, 897 k training interactions, K=5
    shards with compact user tables, batch 30 000, d=16): two epochs of all five shards in one launch against the
    CPU oracle run shard by shard, for both schedules."""
    torch = _torch()
    from ultrare_b200 import kernels as kn, synth
    (u, i, r), _ = synth.ml_like()
    U, I, d, K, batch, epochs = 6040, 3416, 16, 5, 30000, 2
    rs = np.random.RandomState(0)
    groups = np.array_split(rs.permutation(U), K)
    row_of, owner = np.zeros(U, dtype=np.int64), np.zeros(U, dtype=np.int64)
    for g, ids in enumerate(groups):
        row_of[ids], owner[ids] = np.arange(len(ids)), g
    rng = np.random.default_rng(11)
    shards, host = [], []
    for g, ids in enumerate(groups):
        loc = owner[u] == g
        ug, ig, rg = row_of[u[loc]], i[loc], (r[loc] / 5.0).astype(np.float32)
        P0 = rng.standard_normal((len(ids), d), dtype=np.float32)
        Q0 = rng.standard_normal((I, d), dtype=np.float32)
        shards.append(kn.ShardState(kn.pack_interactions(ug, ig, rg, cuda_dev), torch.tensor(P0, device=cuda_dev),
                                    torch.tensor(Q0, device=cuda_dev), epochs, g + 1, 42))
        host.append((ug, ig, rg, P0, Q0))
    assert sum(len(h[0]) for h in host) == len(u) > 890_000
    sb = kn.ShardBatch(shards, d, batch, mode=mode)
    assert sb.mode == mode
    sb.train()
    losses = sb.train_losses()
    for g in range(K):
        ug, ig, rg, P0, Q0 = host[g]
        perms = [omf.feistel_perm(len(ug), omf.perm_key(42, g + 1, ep)) for ep in range(epochs)]
        P, Q, bP, bQ, ls = omf.mf_train(P0, Q0, ug, ig, rg, perms, batch, epochs)
        np.testing.assert_allclose(losses[g], ls, rtol=1e-5)
        assert np.abs(shards[g].P.cpu().numpy() - P).max() < 2e-4
        assert np.abs(shards[g].Q.cpu().numpy() - Q).max() < 2e-4
        assert np.abs(shards[g].bufP.cpu().numpy() - bP).max() < 4e-3      # momenta reach ~60: 6e-5 relative


def test_mf_c3_share_shape_vs_oracle(cuda_dev):
    """One GPU's share of BASELINE config C3 at its real shape (ml-20m: one of 8 shards -- 17 311 users x 26 744 items,
    2.5 M interactions, d = 64, batch 30 000): the wide-row OWNER schedule (four chunks per lane, rows resident in shared
    memory only with the row-balancing item order) against the CPU oracle, two epochs = 168 global steps."""
    torch = _torch()
    from ultrare_b200 import kernels as kn, synth
    U, I, n, d, batch, epochs = 138_493 // 8, 26_744, 20_000_263 // 8, 64, 30000, 2
    rec = synth.device_interactions(U, I, n, cuda_dev, seed=synth.SEED + 3)
    h = rec.cpu().numpy()
    ug, ig, rg = h[:, 0].astype(np.int64), h[:, 1].astype(np.int64), h[:, 2].copy().view(np.float32)
    assert ug.max() < U and ig.max() < I and len(ug) == n
    rng = np.random.default_rng(5)
    P0 = (0.1 * rng.standard_normal((U, d))).astype(np.float32)
    Q0 = (0.1 * rng.standard_normal((I, d))).astype(np.float32)
    sh = kn.ShardState(rec, torch.tensor(P0, device=cuda_dev), torch.tensor(Q0, device=cuda_dev), epochs, 4, 42)
    sb = kn.ShardBatch([sh], d, batch, mode="owner")
    assert sb.mode == "owner"
    sb.train()
    losses = sb.train_losses()[0]
    perms = [omf.feistel_perm(n, omf.perm_key(42, 4, ep)) for ep in range(epochs)]
    P, Q, bP, bQ, ls = omf.mf_train(P0, Q0, ug, ig, rg, perms, batch, epochs)
    np.testing.assert_allclose(losses, ls, rtol=1e-5)
    assert np.abs(sh.P.cpu().numpy() - P).max() < 1e-4 and np.abs(sh.Q.cpu().numpy() - Q).max() < 1e-4
    assert np.abs(sh.bufP.cpu().numpy() - bP).max() < 2e-3 and np.abs(sh.bufQ.cpu().numpy() - bQ).max() < 2e-3


@pytest.mark.parametrize("contiguous", [True, False])
def test_user_segments_on_device_equal_host_segments(cuda_dev, contiguous):
    """Per-user test segments built on the device (stable sort + run starts, padded with empty segments) give the
    same ranking metrics as the host segmentation of utils.py:151-161, for user-sorted and for shuffled rows, with
    users that have no test rows and an id bound larger than the largest id; empty input."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    rng = np.random.default_rng(21)
    n, n_user = 5000, 400
    u = rng.integers(0, 300, n)                     # users 300..399 have no rows
    if contiguous:
        u = np.sort(u)
    i = rng.integers(0, 50, n)
    r = rng.integers(1, 6, n).astype(np.float32) / 5
    inter = kn.pack_interactions(u, i, r, cuda_dev)
    score = torch.tensor(np.round(rng.random(n), 2).astype(np.float32), device=cuda_dev)      # ties on purpose
    order_h, seg_h = kn.user_segments(u)
    assert (order_h is None) == contiguous
    want = kn.rank_metrics(inter, score, kn.upload_array(seg_h, cuda_dev),
                           None if order_h is None else kn.upload_array(order_h, cuda_dev)).cpu().numpy()
    order_d, seg_d = kn.user_segments_device(inter, n_user)
    assert seg_d.shape[0] == n_user + 1 and int(seg_d[0]) == 0 and int(seg_d[-1]) == n
    # bit-exact: the stable argsort by user id and its CSR offsets (hand-written radix sort, no library kernel)
    assert np.array_equal(order_d.cpu().numpy(), np.argsort(u, kind="stable"))
    assert np.array_equal(seg_d.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(u, minlength=n_user))]))
    got = kn.rank_metrics(inter, score, seg_d, order_d).cpu().numpy()
    assert got[2] == want[2] == len(np.unique(u))
    np.testing.assert_allclose(got[:2], want[:2], rtol=1e-12)
    o, s = kn.user_segments_device(inter[:0], n_user)
    assert o is None and int(s.max()) == 0


def test_download_many_and_pinned_upload_round_trip(cuda_dev):
    """download_many == .cpu() of every tensor (mixed dtypes, odd sizes, an empty tensor); records uploaded from a
    page-locked source array (no staging copy) equal the records of the same pageable array."""
    torch = _torch()
    from ultrare_b200 import kernels as kn
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    ts = [torch.randn((37, 16), device=cuda_dev, generator=g), torch.arange(1001, device=cuda_dev, dtype=torch.int64),
          torch.zeros((0, 4), device=cuda_dev), torch.randn(5, device=cuda_dev, generator=g).double()]
    for got, t in zip(kn.download_many(ts), ts):
        assert got.dtype == t.cpu().numpy().dtype and np.array_equal(got, t.cpu().numpy())
    rng = np.random.default_rng(8)
    n = 70000
    raw = np.stack([rng.integers(0, 900, n), rng.integers(0, 500, n), rng.integers(1, 6, n) / 5.0]).astype(np.float64)
    pinned = kn.pinned_copy(raw)
    assert torch.from_numpy(pinned).is_pinned() and np.array_equal(pinned, raw)
    a = kn.upload_interactions(raw, cuda_dev)
    b = kn.upload_interactions(pinned, cuda_dev)
    assert torch.equal(a, b)

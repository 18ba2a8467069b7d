"""CPU: the oracle restatements against the fixtures produced by the reference itself."""
import hashlib

import numpy as np
import pytest

from conftest import init_weights, load_gold
from oracle import evalm, mf as omf, ot as oot, sisa as osisa


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_feistel_is_a_permutation():
    for n in (1, 2, 3, 5, 64, 1000, 28361, 65537):
        p = omf.feistel_perm(n, omf.perm_key(42, 3, 7))
        assert np.array_equal(np.sort(p), np.arange(n))
    a = omf.feistel_perm(1000, omf.perm_key(42, 0, 0))
    b = omf.feistel_perm(1000, omf.perm_key(42, 0, 1))
    assert (a != b).mean() > 0.9


def test_known_answers_appendix_c():
    """pure functions of config.py:47-49 / utils.py:632 / read.py:29-30."""
    d2 = osisa.deletion_set(6040, 2)
    assert len(d2) == 120 and d2[:10].tolist() == [206, 4945, 4481, 5658, 842, 1580, 214, 4227, 3645, 5885]
    assert sha(d2.astype(np.int64)) == "0157347ffc006f62"
    d5 = osisa.deletion_set(6040, 5)
    assert len(d5) == 302 and sha(d5.astype(np.int64)) == "2bbb3e342af53763"
    np.random.seed(0)
    np.random.choice(6040, 120, replace=False)
    assert np.random.choice(6040, 5, replace=False).tolist() == [2680, 1199, 926, 4420, 1670]
    g = osisa.uniform_groups(6040, 5)
    assert g[0][:8] == [206, 4945, 4481, 5658, 842, 1580, 214, 4227]


def test_mf_train_matches_reference(toy):
    """oracle.mf.mf_train == reference baseTrain + SGD/StepLR (utils.py:46-111, scratch.py:65-80)."""
    z = load_gold("toy_train.npz")
    u, i, r = toy["train"]
    r32 = (r / 5.0).astype(np.float32)
    P0, Q0 = init_weights(int(z["weight_seed"]))
    assert sha(P0) == str(z["P0_sha"]) and sha(Q0) == str(z["Q0_sha"])
    epochs = int(z["epochs"])
    perms = [omf.feistel_perm(len(u), omf.perm_key(int(z["perm_seed"]), 0, ep)) for ep in range(epochs)]
    assert sha(perms[0].astype(np.int64)) == str(z["perm0_sha"])
    P, Q, _, _, losses = omf.mf_train(P0, Q0, u, i, r32, perms, int(z["batch"]), epochs)
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-5)
    assert np.abs(P - z["P_final"]).max() < 1e-4 and np.abs(Q - z["Q_final"]).max() < 1e-4


def _steplr_inputs(toy, z):
    n = int(z["n"])
    u, i, r = (a[:n] for a in toy["train"])
    r32 = (r / 5.0).astype(np.float32)
    P0, Q0 = init_weights(int(z["weight_seed"]), int(z["n_user"]), int(z["n_item"]))
    perms = [omf.feistel_perm(n, omf.perm_key(int(z["perm_seed"]), int(z["shard_id"]), ep)) for ep in range(int(z["epochs"]))]
    return u, i, r32, P0, Q0, perms


def test_mf_train_across_two_steplr_boundaries_matches_reference(toy):
    """101 epochs: StepLR(step_size=50, gamma=0.95) fires after epochs 50 and 100 (scratch.py:69,80).  The fixture is
    the reference's own baseTrain + scheduler; a constant learning rate would miss it by 1.4 % in the last loss."""
    z = load_gold("toy_steplr.npz")
    u, i, r32, P0, Q0, perms = _steplr_inputs(toy, z)
    assert omf.lr_at_epoch(1e-3, 0.95, 49) == 1e-3 and abs(omf.lr_at_epoch(1e-3, 0.95, 100) - 1e-3 * 0.95 ** 2) < 1e-12
    P, Q, _, _, losses = omf.mf_train(P0, Q0, u, i, r32, perms, int(z["batch"]), int(z["epochs"]))
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-5)
    assert np.abs(P - z["P_final"]).max() < 1e-4 and np.abs(Q - z["Q_final"]).max() < 1e-4
    _, _, _, _, flat = omf.mf_train(P0, Q0, u, i, r32, perms, int(z["batch"]), int(z["epochs"]), lr_decay=1.0)
    assert abs(flat[-1] - z["losses"][-1]) / z["losses"][-1] > 5e-3          # the fixture is sensitive to the decay


def test_mf_train_torch_port_matches_numpy(toy):
    import torch
    u, i, r = toy["train"]
    r32 = (r / 5.0).astype(np.float32)
    P0, Q0 = init_weights(5)
    perm = omf.feistel_perm(len(u), 99)
    P, Q, bP, bQ = P0.copy(), Q0.copy(), np.zeros_like(P0), np.zeros_like(Q0)
    l1, _ = omf.mf_train_epoch(P, Q, bP, bQ, u, i, r32, perm, 3000, 1e-3, 0.1, 0.9, 0)
    tP, tQ = torch.tensor(P0), torch.tensor(Q0)
    tbP, tbQ = torch.zeros_like(tP), torch.zeros_like(tQ)
    l2, _ = omf.mf_train_epoch_torch(tP, tQ, tbP, tbQ, torch.tensor(u), torch.tensor(i), torch.tensor(r32),
                                     torch.tensor(perm), 3000, 1e-3, 0.1, 0.9, 0)
    assert abs(l1 - l2) / l1 < 1e-5
    assert np.abs(tP.numpy() - P).max() < 1e-4


def test_eval_matches_reference(toy):
    """oracle.evalm.base_test == reference baseTest (utils.py:115-187) on the trained toy model."""
    z = load_gold("toy_train.npz")
    u, i, r = toy["test"]
    r32 = (r / 5.0).astype(np.float32)
    score = evalm.ensemble_score([z["P_final"]], [z["Q_final"]], u, i)
    assert np.abs(score - z["test_score"]).max() < 1e-5
    rmse = float(np.sqrt(evalm.sse(score, r32) / len(u)))
    assert abs(rmse - float(z["test_rmse"])) / float(z["test_rmse"]) < 1e-6
    # reference tie order == this host's default argsort
    nd_ref, hr_ref = evalm.rank_metrics(u, r32, z["test_score"], kind=None)
    assert abs(hr_ref - float(z["test_hr"])) < 1e-12
    assert abs(nd_ref - float(z["test_ndcg_ref"])) < 1e-9
    # documented kernel rule (stable) -- HR identical, NDCG within the reference's own tie spread
    nd_st, hr_st = evalm.rank_metrics(u, r32, z["test_score"], kind="stable")
    assert abs(hr_st - hr_ref) < 1e-12
    assert abs(nd_st - nd_ref) / nd_ref < 2e-2


def test_read_rating_and_routing_match_reference(toy):
    """Shard materialisation (read.py:9-70) and routing/merge (sisa.py:52-58,76-81,107-113)."""
    z = load_gold("toy_sisa.npz")
    K = int(z["K"])
    del_user = z["del_user"]
    assert np.array_equal(del_user, osisa.deletion_set(1508, 2))
    u, i, r = toy["train"]
    groups0 = osisa.uniform_groups(1508, K)
    for is_del, tag in ((False, "learn"), (True, "unlearn")):
        lists, idx = osisa.read_rating(u, i, r, 1508, 5, del_user if is_del else (), K, groups0, "a")
        for s in range(K):
            assert np.array_equal(np.asarray(idx[s]), z[f"group{s}"])
            assert lists[s].shape[1] == int(z[f"{tag}_train_n{s}"])
            assert sha(lists[s]) == str(z[f"{tag}_train_sha{s}"])
    assert sorted(osisa.route_deletions(idx, del_user)) == z["retrain_gid"].tolist()


def test_ot_cluster_matches_reference():
    """oracle.ot.ot_cluster == reference ot_cluster (utils.py:628-656) given the same exact-LP plan solver."""
    z = load_gold("ot_cluster.npz")
    np.random.seed(int(z["np_seed"]))
    inertia, label, _, _ = oot.ot_cluster(z["X"], int(z["k"]))
    assert np.array_equal(label, z["label"])
    assert abs(float(inertia) - float(z["inertia"])) / float(z["inertia"]) < 1e-6
    assert np.bincount(label).tolist() == [150] * 4


def test_emd_plan_is_integral_and_balanced():
    rng = np.random.default_rng(3)
    n, k = 300, 5
    M = rng.random((n, k)) * 10
    G = oot.emd_lp(np.ones(n) / n, np.ones(k) / k, M)
    assert np.allclose(G.sum(1), 1 / n) and np.allclose(G.sum(0), 1 / k)
    assert (G > 1e-12).sum() == n                      # Appendix C: integral when k | n
    assert np.bincount(oot.assign(G)).tolist() == [n // k] * k


def test_emd_stand_in_agrees_with_an_independent_exact_solver():
    """ot.emd (POT, absent here) is an exact solver; its stand-in is a dual-simplex LP.  An independent exact method
    -- the assignment problem on centroid columns replicated n/k times (scipy's modified Jonker-Volgenant) -- reaches
    the same vertex on squared-distance costs at the reference's shapes: the optimum is unique there, so ANY exact
    solver, POT's network simplex included, returns this plan.  Also on the reference's own toy embedding (golden)."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(12)
    cases = [(rng.standard_normal((600, 16)), 5), (rng.standard_normal((960, 8)), 32)]
    z = load_gold("ot_cluster.npz")
    cases.append((np.asarray(z["X"], dtype=np.float64), int(z["k"])))
    for X, k in cases:
        n = X.shape[0]
        assert n % k == 0
        M = oot.cost_matrix(X, X[rng.choice(n, k, replace=False)])
        G = oot.emd_lp(np.ones(n) / n, np.ones(k) / k, M)
        rows, cols = linear_sum_assignment(np.repeat(M, n // k, axis=1))
        label = np.empty(n, dtype=np.int64)
        label[rows] = cols // (n // k)
        assert np.array_equal(oot.assign(G), label)
        assert abs((G * M).sum() - M[np.arange(n), label].sum() / n) <= 1e-9 * abs((G * M).sum())


def test_sinkhorn_converges_to_emd_labels():
    rng = np.random.default_rng(4)
    n, k, d = 400, 4, 8
    X = rng.standard_normal((n, d))
    Cc = X[rng.choice(n, k, replace=False)]
    M = oot.cost_matrix(X, Cc)
    G = oot.emd_lp(np.ones(n) / n, np.ones(k) / k, M)
    mean = M.mean()
    P, f, g, err = oot.sinkhorn_log(M, [(mean * e, 200) for e in (0.5, 0.1, 0.02)])
    assert np.allclose(P.sum(1), 1 / n)
    assert err < 1e-3 / k
    assert (oot.assign(P) == oot.assign(G)).mean() > 0.97


def test_decay_matrix_power_matches_stepping():
    lr, wd, mu = 1e-3, 0.1, 0.9
    w, b = np.float64(0.7), np.float64(-0.2)
    for _ in range(37):
        b = mu * b + wd * w
        w = w - lr * b
    Mn = omf.decay_matrix_power(lr, wd, mu, 37)
    w2, b2 = Mn @ np.array([0.7, -0.2])
    assert abs(w - w2) < 1e-12 and abs(b - b2) < 1e-12

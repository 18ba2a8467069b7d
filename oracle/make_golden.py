"""Write tests/golden/*.npz by executing the reference itself (dev container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden
Needs /root/reference (read-only); see oracle/ref_shim.py for the shims.  The
fixtures are small (toy data: 1508 users x 2071 items, 28361 / 7133 rows) and are
committed so that the CPU and GPU test suites never need /root/reference.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle import mf as omf            # noqa: E402
from oracle import ref_shim, sisa as osisa   # noqa: E402

N_USER, N_ITEM, K_DIM, BATCH = 1508, 2071, 16, 3000
SEED = 42


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_toy():
    import pandas as pd
    tr = pd.read_csv(os.path.join(ref_shim.REFERENCE, "data/toy/0_train.csv"), header=None)
    te = pd.read_csv(os.path.join(ref_shim.REFERENCE, "data/toy/0_test.csv"), header=None)
    return tr, te


def arr3(df, max_rating=5):
    """What readRating hands to RatingData for one group (read.py:64-68)."""
    a = df.values.T.astype(np.float64).copy()
    a[2] /= max_rating
    return a


def init_weights(seed, n_user=N_USER, n_item=N_ITEM, k=K_DIM):
    """N(0,1) fp32 tables (utils.py:27,38-40) under the harness's own seed discipline (H8)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_user, k), dtype=np.float32),
            rng.standard_normal((n_item, k), dtype=np.float32))


def gold_data():
    tr, te = load_toy()
    np.savez_compressed(
        os.path.join(GOLD, "toy_data.npz"),
        train_u=tr[0].values.astype(np.int16), train_i=tr[1].values.astype(np.int16),
        train_r2=(tr[2].values * 2).astype(np.uint8),
        test_u=te[0].values.astype(np.int16), test_i=te[1].values.astype(np.int16),
        test_r2=(te[2].values * 2).astype(np.uint8))
    return tr, te


def gold_train(tr, te):
    """Reference baseTrain (+SGD/StepLR of scratch.py) from injected weights and Feistel perms."""
    epochs = 3
    P0, Q0 = init_weights(1234)
    n = len(tr)
    perms = [omf.feistel_perm(n, omf.perm_key(SEED, 0, ep)) for ep in range(epochs)]
    model, losses = ref_shim.train_injected(arr3(tr), N_USER, N_ITEM, K_DIM, P0, Q0, perms, BATCH, epochs)
    P = model.user_mat.weight.detach().numpy()
    Q = model.item_mat.weight.detach().numpy()
    rmse, ndcg, hr = ref_shim.base_test(arr3(te), [model], BATCH)
    import torch
    a = arr3(te)
    with torch.no_grad():
        score = model(torch.as_tensor(a[0].astype(np.int64)), torch.as_tensor(a[1].astype(np.int64))).numpy()
    np.savez_compressed(
        os.path.join(GOLD, "toy_train.npz"),
        weight_seed=1234, perm_seed=SEED, epochs=epochs, batch=BATCH,
        P0_sha=sha(P0), Q0_sha=sha(Q0), perm0_sha=sha(perms[0].astype(np.int64)),
        losses=np.array(losses), P_final=P, Q_final=Q,
        test_rmse=rmse, test_ndcg_ref=ndcg, test_hr=hr, test_score=score.astype(np.float32))
    print("train losses", losses, "rmse/ndcg/hr", rmse, ndcg, hr)


def gold_steplr(tr):
    """Reference baseTrain + StepLR(50, 0.95) (scratch.py:69,80) over 101 epochs -- two decay boundaries -- on the
    first 2400 toy rows, batch 800, injected weights + Feistel visiting orders."""
    epochs, n, batch = 101, 2400, 800
    sub = arr3(tr.iloc[:n])
    U, I = int(sub[0].max()) + 1, int(sub[1].max()) + 1
    P0, Q0 = init_weights(4321, U, I)
    perms = [omf.feistel_perm(n, omf.perm_key(SEED, 3, ep)) for ep in range(epochs)]
    model, losses = ref_shim.train_injected(sub, U, I, K_DIM, P0, Q0, perms, batch, epochs)
    np.savez_compressed(os.path.join(GOLD, "toy_steplr.npz"), weight_seed=4321, perm_seed=SEED, shard_id=3, epochs=epochs,
                        batch=batch, n=n, n_user=U, n_item=I, losses=np.array(losses),
                        P_final=model.user_mat.weight.detach().numpy(), Q_final=model.item_mat.weight.detach().numpy())
    # sensitivity: the same run with a constant learning rate must be visibly different
    Pc, Qc, _, _, lc = omf.mf_train(P0, Q0, sub[0].astype(np.int64), sub[1].astype(np.int64), sub[2].astype(np.float32),
                                    perms, batch, epochs, lr_decay=1.0)
    Pd, Qd, _, _, ld = omf.mf_train(P0, Q0, sub[0].astype(np.int64), sub[1].astype(np.int64), sub[2].astype(np.float32),
                                    perms, batch, epochs)
    print("steplr: reference last loss", losses[-1], "oracle", ld[-1], "constant-lr oracle", lc[-1],
          "max|dP| decay vs constant", np.abs(Pc - Pd).max(), "oracle vs reference",
          np.abs(Pd - model.user_mat.weight.detach().numpy()).max())


def gold_sisa(tr, te):
    """Reference Sisa.learn + Sisa.unlearn on toy, K=3, 2 epochs, injected init + perms."""
    import torch
    ns = ref_shim.load()
    K, epochs = 3, 2
    del_user = osisa.deletion_set(N_USER, 2)                      # 30 users
    groups0 = osisa.uniform_groups(N_USER, K)

    # the reference's readRating wants a path containing 'ml1m' (read.py:41-44)
    d = os.path.join(ns.root, "data", "ml1m")
    tr.to_csv(os.path.join(d, "squ0_train.csv"), header=False, index=False)
    te.to_csv(os.path.join(d, "squ0_test.csv"), header=False, index=False)
    import contextlib, io

    def read(is_del):
        with contextlib.redirect_stdout(io.StringIO()):          # read.py:57-58 prints whole groups
            trr, idx = ns.read.readRating(os.path.join(d, "squ0_train.csv"), N_USER, 5,
                                          list(del_user) if is_del else [], [], K, groups0, "a")
            ter, _ = ns.read.readRating(os.path.join(d, "squ0_test.csv"), N_USER, 5, [], [], K, idx)
        return trr, idx, ter

    out = {"del_user": del_user.astype(np.int64), "K": K, "epochs": epochs}
    weights = [init_weights(100 + s) for s in range(K)]            # shard id s -> init seed 100+s
    state = {"next": 0}

    def fake_MF(n_user, n_item, k):
        s = state["next"]
        P0, Q0 = weights[s]
        return ref_shim.make_model(n_user, n_item, k, P0, Q0)

    ns.scratch.MF = fake_MF                                         # module attribute of the temp copy
    param = ref_shim.Param(N_USER, N_ITEM, epochs, BATCH)

    def loaders(trr, ter, phase):
        tl, sl = [], []
        for s in range(K):
            n = trr[s].shape[1]
            perms = [omf.feistel_perm(n, omf.perm_key(SEED + phase, s, ep)) for ep in range(epochs)]
            tl.append(ref_shim.PermLoader(trr[s], BATCH, perms))
            sl.append(ref_shim.PermLoader(ter[s], BATCH, None))
        total = ref_shim.PermLoader(np.hstack(ter), BATCH, None)    # config.py:144-151
        return tl, sl, total

    save_dir = os.path.join(ns.root, "result", "sisa_gold")
    os.makedirs(save_dir, exist_ok=True)

    # ---- learn (config.py:164-166 -> sisa.py:25-63)
    trr, idx, ter = read(False)
    for s in range(K):
        out[f"group{s}"] = np.asarray(idx[s], dtype=np.int64)
        out[f"learn_train_n{s}"] = trr[s].shape[1]
        out[f"learn_train_sha{s}"] = sha(trr[s])
        out[f"test_n{s}"] = ter[s].shape[1]
    tl, sl, total = loaders(trr, ter, 0)
    sisa = ns.sisa.Sisa(param, "mf", K, idx)
    orig_train = ns.scratch.Scratch.train

    def train_hook(self, train_data, test_data, test_total=[], verbose=1, save_dir="", id=0, given_model=""):
        state["next"] = id - 1
        return orig_train(self, train_data, test_data, test_total, verbose, save_dir, id, given_model)

    ns.scratch.Scratch.train = train_hook
    with contextlib.redirect_stdout(io.StringIO()):
        model_list = sisa.learn(tl, sl, total, 0, save_dir)
    log0 = np.load(save_dir + "/log0.npy", allow_pickle=True).item()
    out["learn_log0"] = np.array([log0["total_rmse"], log0["total_ndcg"], log0["total_hr"]])
    out["learn_merged"] = model_list[0].user_mat.weight.detach().numpy().copy()
    for s in range(K):
        out[f"learn_Q{s}"] = model_list[s].item_mat.weight.detach().numpy().copy()
    for key in ("train_loss", "test_rmse", "test_ndcg", "test_hr", "total_rmse", "total_ndcg", "total_hr"):
        out["learn_log_" + key] = np.array(sisa.log[key], dtype=np.float64)   # accumulates over shards (A14)

    # ---- unlearn (config.py:168-172 -> sisa.py:66-118)
    trr, idx2, ter = read(True)
    assert all(list(a) == list(b) for a, b in zip(idx, idx2))
    for s in range(K):
        out[f"unlearn_train_n{s}"] = trr[s].shape[1]
        out[f"unlearn_train_sha{s}"] = sha(trr[s])
    tl, sl, total = loaders(trr, ter, 1)
    before = list(model_list)
    sisa2 = ns.sisa.Sisa(param, "mf", K, idx2)
    with contextlib.redirect_stdout(io.StringIO()):
        model_list2 = sisa2.unlearn(list(model_list), tl, sl, total, list(del_user), 0, save_dir)
    retrained = [s for s in range(K) if model_list2[s] is not before[s]]
    out["retrain_gid"] = np.array(retrained, dtype=np.int64)
    log0 = np.load(save_dir + "/log0.npy", allow_pickle=True).item()
    out["unlearn_log0"] = np.array([log0["total_rmse"], log0["total_ndcg"], log0["total_hr"]])
    out["unlearn_merged"] = model_list2[0].user_mat.weight.detach().numpy().copy()
    for s in range(K):
        out[f"unlearn_Q{s}"] = model_list2[s].item_mat.weight.detach().numpy().copy()
    out["unlearn_log_train_loss"] = np.array(sisa2.log["train_loss"], dtype=np.float64)
    ns.scratch.Scratch.train = orig_train
    np.savez_compressed(os.path.join(GOLD, "toy_sisa.npz"), **out)
    print("sisa learn log0", out["learn_log0"], "unlearn log0", out["unlearn_log0"], "retrained", retrained)


def gold_ot():
    """Reference ot_cluster (utils.py:628-656) with the exact-LP ot.emd stand-in."""
    import contextlib, io
    ns = ref_shim.load()
    rng = np.random.default_rng(7)
    k, n, d = 4, 600, 16
    centers = rng.standard_normal((k, d)) * 2.0
    X = (centers[rng.integers(0, k, n)] + rng.standard_normal((n, d))).astype(np.float32)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        inertia, label = ns.utils.ot_cluster(X, k)
    np.savez_compressed(os.path.join(GOLD, "ot_cluster.npz"), X=X, k=k, np_seed=0,
                        inertia=np.float64(inertia), label=label.astype(np.int64))
    print("ot_cluster inertia", inertia, "sizes", np.bincount(label))


def main():
    os.makedirs(GOLD, exist_ok=True)
    tr, te = gold_data()
    gold_train(tr, te)
    gold_sisa(tr, te)
    gold_steplr(tr)
    gold_ot()


if __name__ == "__main__":
    main()

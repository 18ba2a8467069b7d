"""GPU, 2 ranks over NCCL (skipped on a 1-GPU box): the multi-GPU SISA and Sinkhorn paths give the
single-GPU results."""
import os
import socket

import numpy as np
import pytest
from conftest import collect_results

pytestmark = pytest.mark.gpu
N_USER, N_ITEM, BATCH, SEED = 1508, 2071, 3000, 42


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "toy_data.npz"))
    tr = (z["train_u"].astype(np.int64), z["train_i"].astype(np.int64), z["train_r2"] / 2.0)
    te = (z["test_u"].astype(np.int64), z["test_i"].astype(np.int64), z["test_r2"] / 2.0)
    return tr, te


class Param:
    def __init__(self, epochs):
        self.n_user, self.n_item, self.k, self.lam = N_USER, N_ITEM, 16, 0.1
        self.seed, self.lr, self.lr_decay, self.momentum = SEED, 0.001, 0.95, 0.9
        self.epochs, self.batch = epochs, BATCH


def _sisa_pass(dist_obj, eval_sharded=None):
    """learn + unlearn on toy, K=4, host-seeded init (identical weights on every rank layout).
    eval_sharded: None = final evaluation with the test rows sharded over the ranks, False = every rank scores
    the whole merged test set and the partial scores are all-reduced."""
    import pandas as pd
    import torch
    from oracle import sisa as osisa
    from ultrare_b200.method.sisa import Sisa
    from ultrare_b200.read import RatingData, loadData, readRating
    K, E = 4, 2
    tr, te = _toy()
    dtr = pd.DataFrame({0: tr[0], 1: tr[1], 2: tr[2]})
    dte = pd.DataFrame({0: te[0], 1: te[1], 2: te[2]})
    g0 = osisa.uniform_groups(N_USER, K)
    del_user = list(osisa.deletion_set(N_USER, 5))
    out = {}
    models = None
    for phase, dels in (("learn", []), ("unlearn", del_user)):
        trr, idx = readRating(dtr, N_USER, 5, dels, [], K, g0, 'a')
        ter, _ = readRating(dte, N_USER, 5, [], [], K, idx)
        tl = [loadData(RatingData(a), BATCH, 1, True) for a in trr]
        sl = [loadData(RatingData(a), BATCH, 1, False) for a in ter]
        total = loadData(RatingData(np.hstack(ter)), BATCH, 1, False)
        s = Sisa(Param(E), 'mf', K, idx)
        s.init_on_device = False
        s.epoch_eval = 'none'
        s.eval_sharded = eval_sharded
        if dist_obj is not None:
            s.dist = dist_obj
        if phase == "learn":
            models = s.learn(tl, sl, total, 0, '')
        else:
            models = s.unlearn(models, tl, sl, total, del_user, 0, '')
            out["retrain_gid"] = sorted(s.retrain_gid)
        out[phase + "_merged"] = models[0].user_mat.weight.data.cpu().numpy()
        out[phase + "_log0"] = [s.final_log[k] for k in ('total_rmse', 'total_ndcg', 'total_hr')]
    torch.cuda.synchronize()
    return out


def _optimistic_pass(dist_obj):
    """Device-initialised batches (kernels.ArenaShardBatch) on several GPUs: a pass queued on the remembered plan
    equals the pass that waited for its plan, and a remembered plan that does not cover ONE rank's batch makes every
    rank repeat the pass (the flag travels with the all-reduced metric sums)."""
    import pandas as pd
    import torch
    from oracle import sisa as osisa
    from ultrare_b200 import kernels as kn
    from ultrare_b200.method.sisa import Sisa
    from ultrare_b200.read import RatingData, loadData, readRating
    K, E = 4, 2
    tr, te = _toy()
    dtr = pd.DataFrame({0: tr[0], 1: tr[1], 2: tr[2]})
    dte = pd.DataFrame({0: te[0], 1: te[1], 2: te[2]})
    g0 = osisa.uniform_groups(N_USER, K)
    trr, idx = readRating(dtr, N_USER, 5, [], [], K, g0, 'a')
    ter, _ = readRating(dte, N_USER, 5, [], [], K, idx)
    tl = [loadData(RatingData(a), BATCH, 1, True) for a in trr]
    sl = [loadData(RatingData(a), BATCH, 1, False) for a in ter]
    total = loadData(RatingData(np.hstack(ter)), BATCH, 1, False)
    out = []
    for what in ("wait", "hit", "miss"):
        if what == "wait":
            kn._PLAN_HINTS.clear()
        if what == "miss" and dist_obj.rank == 1:
            for key, v in list(kn._PLAN_HINTS.items()):
                kn._PLAN_HINTS[key] = (max(1, v[0] // 2), max(16, v[1] // 2), v[2], v[3])
        s = Sisa(Param(E), 'mf', K, idx)
        s.epoch_eval = 'none'
        s.dist = dist_obj
        models = s.learn(tl, sl, total, 0, '')
        sb = s._last_batch
        out.append(dict(what=what, optimistic=bool(sb is not None and sb.optimistic), mode=None if sb is None else sb.mode,
                        merged=models[0].user_mat.weight.data.cpu().numpy(),
                        log0=[s.final_log[k] for k in ('total_rmse', 'total_ndcg', 'total_hr')]))
    torch.cuda.synchronize()
    return out


def _ot_pass(dist_obj):
    from ultrare_b200.method.utils import ot_cluster_device
    rng = np.random.default_rng(3)
    n, k, d = 4000, 8, 16
    X = rng.standard_normal((n, d), dtype=np.float32)
    c0 = X[:k].copy()
    if dist_obj is None or dist_obj.world == 1:
        # the row-sharded path skips the balanced rounding (the cost rows live on different GPUs): same here
        inertia, label, cen, it = ot_cluster_device(X, k, centroid0=c0, balance=False)
        return dict(inertia=float(inertia), label=label, it=it)
    import torch
    from ultrare_b200 import kernels as kn
    lo, hi = dist_obj.row_block(n)
    inertia, label, cen, it = ot_cluster_device(X[lo:hi], k, centroid0=c0, dist=dist_obj)
    out = dict(inertia=float(inertia), label=label, it=it, lo=lo, hi=hi, peer=kn._peer_exchange(dist_obj, torch.device("cuda", dist_obj.rank)) is not None)
    # the fused peer-memory Sinkhorn against the NCCL all-reduce loop on the same cost rows: same potentials
    dev = torch.device("cuda", dist_obj.rank)
    M = kn.cost_matrix(torch.tensor(X[lo:hi], device=dev), torch.tensor(c0, device=dev))
    sched = [(8.0, 15), (2.0, 25)]
    g_peer = kn.sinkhorn_sharded(M, k, sched, dist_obj, n)
    kn.PEER_SINKHORN = False
    g_nccl = kn.sinkhorn_sharded(M, k, sched, dist_obj, n)
    kn.PEER_SINKHORN = True
    out["g_peer"], out["g_nccl"] = g_peer.cpu().numpy(), g_nccl.cpu().numpy()
    return out


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ultrare_b200 import dist as udist
    d = udist.init_from_env(backend="nccl")
    res = dict(sisa=_sisa_pass(d), sisa_whole=_sisa_pass(d, eval_sharded=False), ot=_ot_pass(d), opt=_optimistic_pass(d))
    d.barrier()
    q.put((rank, res))
    d.td.destroy_process_group()


def test_two_rank_sisa_and_sinkhorn_equal_single_gpu(cuda_dev):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = dict(sisa=_sisa_pass(None), ot=_ot_pass(None))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = collect_results(procs, q, 2, timeout=300)      # a rank that raises ends the test at once
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        for variant in ("sisa", "sisa_whole"):          # both multi-GPU evaluation paths against one GPU
            s, ref = results[r][variant], single["sisa"]
            assert s["retrain_gid"] == ref["retrain_gid"]
            for phase in ("learn", "unlearn"):
                assert np.abs(s[phase + "_merged"] - ref[phase + "_merged"]).max() < 1e-4
                np.testing.assert_allclose(s[phase + "_log0"], ref[phase + "_log0"], rtol=1e-3)
        np.testing.assert_allclose(results[r]["sisa"]["unlearn_log0"], results[r]["sisa_whole"]["unlearn_log0"], rtol=1e-5)
    for r in range(2):
        wait, hit, miss = results[r]["opt"]
        assert wait["mode"] == hit["mode"] == "owner"
        assert not wait["optimistic"] and hit["optimistic"] and not miss["optimistic"]      # the miss was repeated
        for other in (hit, miss):
            assert np.array_equal(other["merged"], wait["merged"])
            np.testing.assert_allclose(other["log0"], wait["log0"], rtol=1e-12)
    for r in range(2):
        assert results[r]["ot"]["peer"], "no peer-mapped symmetric memory between the two GPUs"
        assert np.abs(results[r]["ot"]["g_peer"] - results[r]["ot"]["g_nccl"]).max() < 1e-5
    assert np.array_equal(results[0]["ot"]["g_peer"], results[1]["ot"]["g_peer"])      # bit-identical on every rank
    lab = np.concatenate([results[0]["ot"]["label"], results[1]["ot"]["label"]])
    assert (lab == single["ot"]["label"]).mean() > 0.999
    assert abs(results[0]["ot"]["inertia"] - single["ot"]["inertia"]) / single["ot"]["inertia"] < 1e-5

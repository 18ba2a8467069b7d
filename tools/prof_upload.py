"""Time the host -> device ingest of the bench workload's shards (staging copies, H2D, pack kernel)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ultrare_b200 import kernels as kn
w = bench.host_workload(0, 50)
rs = np.random.RandomState(1)
perm = rs.permutation(w["n_user"])
sp = bench.group_and_split(w, [perm[i::5].tolist() for i in range(5)])
raws = sp["unlearn_train"]
dev = torch.device("cuda:0")
row_of = torch.zeros(w["n_user"], dtype=torch.int32, device=dev)
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = kn.upload_interactions_many(raws, dev, row_of)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: host returns after {1e3 * (t1 - t0):.3f} ms, device done after {1e3 * (t2 - t0):.3f} ms "
          f"({sum(r.shape[1] for r in raws) * 24 / 1e6:.1f} MB)")
# pieces
n = raws[0].shape[1]
st = kn._staging_bytes(24 * n)
dst = st.numpy()[:24 * n].view(np.float64).reshape(3, n)
t0 = time.perf_counter()
for _ in range(20): np.copyto(dst, raws[0])
print("one shard staging copy (%.1f MB): %.3f ms" % (24 * n / 1e6, (time.perf_counter() - t0) / 20 * 1e3))
pool = kn._pack_pool()
stages = [kn._staging_bytes(24 * r.shape[1]) for r in raws]
def fill(j):
    m = raws[j].shape[1]
    np.copyto(stages[j].numpy()[:24 * m].view(np.float64).reshape(3, m), raws[j])
for _ in range(3): list(pool.map(fill, range(5)))
t0 = time.perf_counter()
for _ in range(20): list(pool.map(fill, range(5)))
print("five staging copies on the pool: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))

# ---- GPU timeline of one upload: events around every H2D and pack
import ctypes as C
from ultrare_b200 import _lib
ns = [r.shape[1] for r in raws]; n_tot = sum(ns)
offs = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
stage = kn._staging_bytes(24 * n_tot); stage_f = stage[:24 * n_tot].view(torch.float64)
stage_np = stage.numpy()[:24 * n_tot].view(np.float64)
cols_all = torch.empty(3 * n_tot, dtype=torch.float64, device=dev); out_all = torch.empty((n_tot, 4), dtype=torch.int32, device=dev)
def fill2(j): np.copyto(stage_np[3 * offs[j]:3 * offs[j + 1]].reshape(3, ns[j]), raws[j])
for rep in range(3):
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    host = []
    t0 = time.perf_counter()
    evs[0].record()
    futs = [pool.submit(fill2, j) for j in range(5)]
    for j, f in enumerate(futs):
        f.result(); host.append(time.perf_counter() - t0)
        lo, hi = 3 * int(offs[j]), 3 * int(offs[j + 1])
        cols_all[lo:hi].copy_(stage_f[lo:hi], non_blocking=True)
        evs[1 + 2 * j].record()
        _lib.check(_lib.lib().ure_pack_interactions_f64(C.c_void_p(cols_all.data_ptr() + 8 * lo), ns[j], ns[j], None, 0,
                                                        C.c_void_p(out_all[offs[j]:offs[j + 1]].data_ptr()), None))
        evs[2 + 2 * j].record()
        host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print("host (copy done, shipped) ms:", np.round(np.array(host) * 1e3, 3))
    print("gpu  (h2d done, pack done) ms:", np.round([evs[0].elapsed_time(e) for e in evs[1:]], 3))

print("---- variant: one H2D of everything after all staging copies, per-shard pack kernels")
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter(); e0.record()
    list(pool.map(fill2, range(5)))
    tc = time.perf_counter() - t0
    cols_all.copy_(stage_f, non_blocking=True); e1.record()
    for j in range(5):
        lo = 3 * int(offs[j])
        _lib.check(_lib.lib().ure_pack_interactions_f64(C.c_void_p(cols_all.data_ptr() + 8 * lo), ns[j], ns[j], None, 0,
                                                        C.c_void_p(out_all[offs[j]:offs[j + 1]].data_ptr()), None))
    e2.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"copies done {tc*1e3:.3f} ms, host returns {th*1e3:.3f} ms, h2d done {e0.elapsed_time(e1):.3f} ms, packs done {e0.elapsed_time(e2):.3f} ms")
print("---- variant: per-shard H2D back to back (no kernels in between), then packs")
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter(); e0.record()
    futs = [pool.submit(fill2, j) for j in range(5)]
    for j, f in enumerate(futs):
        f.result()
        lo, hi = 3 * int(offs[j]), 3 * int(offs[j + 1])
        cols_all[lo:hi].copy_(stage_f[lo:hi], non_blocking=True)
    e1.record()
    for j in range(5):
        lo = 3 * int(offs[j])
        _lib.check(_lib.lib().ure_pack_interactions_f64(C.c_void_p(cols_all.data_ptr() + 8 * lo), ns[j], ns[j], None, 0,
                                                        C.c_void_p(out_all[offs[j]:offs[j + 1]].data_ptr()), None))
    e2.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"host returns {th*1e3:.3f} ms, h2d done {e0.elapsed_time(e1):.3f} ms, packs done {e0.elapsed_time(e2):.3f} ms")

print("---- diagnosis: one H2D of everything, (a) right after pool copies, (b) after pool copies + 3 ms sleep, (c) after main-thread copies")
def one(mode):
    torch.cuda.synchronize()
    if mode == "main":
        for j in range(5): fill2(j)
    else:
        list(pool.map(fill2, range(5)))
    if mode == "sleep": time.sleep(0.003)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cols_all.copy_(stage_f, non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for mode in ("pool", "sleep", "main", "pool", "sleep", "main"):
    print(mode, "H2D of 21 MB: %.3f ms" % one(mode))
big = torch.empty(24 * n_tot, dtype=torch.uint8, pin_memory=True)
bf = big.view(torch.float64); bn = big.numpy().view(np.float64)
def fill3(j): np.copyto(bn[3 * offs[j]:3 * offs[j + 1]].reshape(3, ns[j]), raws[j])
for mode in ("fresh pinned buffer, pool copies",) * 3:
    torch.cuda.synchronize(); list(pool.map(fill3, range(5)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cols_all.copy_(bf, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print(mode, "H2D: %.3f ms" % e0.elapsed_time(e1), "pinned?", big.is_pinned(), stage.is_pinned(), stage_f.is_pinned(), "stage bytes", stage.numel())

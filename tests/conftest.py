import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.fixture(scope="session")
def toy():
    """The reference's data/toy/0_{train,test}.csv (fixture written by oracle/make_golden.py)."""
    z = load_gold("toy_data.npz")
    out = {}
    for part in ("train", "test"):
        out[part] = (z[part + "_u"].astype(np.int64), z[part + "_i"].astype(np.int64),
                     z[part + "_r2"].astype(np.float64) / 2.0)
    out["n_user"], out["n_item"], out["k"], out["batch"] = 1508, 2071, 16, 3000
    return out


def init_weights(seed, n_user=1508, n_item=2071, k=16):
    """Same generator as oracle/make_golden.py:init_weights."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_user, k), dtype=np.float32),
            rng.standard_normal((n_item, k), dtype=np.float32))


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def collect_results(procs, q, n, timeout=300.0, poll=1.0):
    """{rank: result} from `n` worker processes that each put (rank, result) on `q`.  A worker that dies first (an
    exception on one rank leaves the others waiting in a collective) ends the wait at once: the survivors are killed
    and the test fails with the exit codes, instead of holding the GPUs until the time-out."""
    import queue
    import time
    results, deadline = {}, time.monotonic() + timeout
    while len(results) < n:
        try:
            rank, res = q.get(timeout=poll)
            results[rank] = res
            continue
        except queue.Empty:
            pass
        dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
        if dead or time.monotonic() > deadline:
            for p in procs:
                if p.is_alive():
                    p.kill()
            for p in procs:
                p.join(10)
            pytest.fail(f"workers failed: exit codes {[p.exitcode for p in procs]}" if dead else
                        f"no result from {n - len(results)} worker(s) after {timeout:.0f} s")
    return results

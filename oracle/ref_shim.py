"""Run the UNMODIFIED reference (ZhangYizhao/UltraRE) in the dev container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
``/root/reference`` exists (the dev container); it is used by
``oracle/make_golden.py`` to write ``tests/golden/*.npz`` and by the
``not gpu`` tests that re-validate the restatements when the reference is
present.  Nothing on the GPU box imports this module.

The reference does not run as shipped; the shims
applied here, none of which edits a reference file:
  A1   del_per is a percentage: InsParam is not used; callers pass del_user from
       oracle.sisa.deletion_set
  A2   param.dis_type = 'nor', param.attr = []
  A6   torch.manual_seed() before anything that draws
  A8   data paths contain 'ml1m' (read.py:41-44)
  --   stub modules ``ot`` (exact LP from oracle.ot.emd_lp standing in for POT 0.9.0)
       and ``matplotlib.pyplot`` (utils.py:4,14 import them at module import)
  --   the tree is copied to a temp dir first because config/group compute
       SAVE_DIR/DATA_DIR next to their own files and /root/reference is read-only
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE = "/root/reference"
_loaded = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "method"))


def load():
    """Import the reference modules from a temp copy; returns a namespace of modules."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE)
    root = tempfile.mkdtemp(prefix="ultrare_ref_")
    for name in ("main.py", "config.py", "group.py", "read.py"):
        shutil.copy(os.path.join(REFERENCE, name), root)
    os.makedirs(os.path.join(root, "method"))
    for name in ("scratch.py", "sisa.py", "utils.py"):
        shutil.copy(os.path.join(REFERENCE, "method", name), os.path.join(root, "method"))
    os.makedirs(os.path.join(root, "data", "ml1m", "val"))
    os.makedirs(os.path.join(root, "result"))

    from oracle import ot as oracle_ot

    ot_stub = types.ModuleType("ot")
    ot_stub.emd = lambda a, b, M, numItermax=100000, **kw: oracle_ot.emd_lp(a, b, M)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("ot", ot_stub)
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("config", "group", "read", "method") or k.startswith("method.")}
    sys.path.insert(0, root)
    try:
        import importlib
        ns = types.SimpleNamespace(root=root)
        ns.utils = importlib.import_module("method.utils")
        ns.scratch = importlib.import_module("method.scratch")
        ns.sisa = importlib.import_module("method.sisa")
        ns.read = importlib.import_module("read")
        ns.group = importlib.import_module("group")
    finally:
        sys.path.remove(root)
        for k in list(sys.modules):
            if k in ("config", "group", "read", "method") or k.startswith("method."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    _loaded["ns"] = ns
    return ns


class Param:
    """The fields Scratch/Sisa read from InsParam (config.py:17-38) + shim A2."""

    def __init__(self, n_user, n_item, epochs, batch, k=16, seed=42):
        self.n_user, self.n_item, self.k = n_user, n_item, k
        self.lam = 0.1
        self.seed = seed
        self.lr, self.lr_decay, self.momentum = 0.001, 0.95, 0.9
        self.epochs = epochs
        self.batch = batch
        self.dis_type, self.attr = "nor", []


class PermLoader:
    """Stands in for DataLoader(RatingData, batch, shuffle) (read.py:108-133).

    ``baseTrain``/``baseTest`` only use ``len(loader.dataset)`` and iteration
    yielding (user int64, item int64, rating fp32) tensors (utils.py:52,58,129).
    Successive ``__iter__`` calls consume ``perms`` in order (one per epoch); with
    ``perms=None`` rows come in file order (shuffle=False).
    """

    def __init__(self, arr3, batch, perms=None):
        import torch
        ns = load()
        self.dataset = ns.read.RatingData(arr3)      # reference casts: read.py:111-113
        self.batch = batch
        self.perms = None if perms is None else [np.asarray(p) for p in perms]
        self.calls = 0
        self._u = torch.tensor(self.dataset.users, dtype=torch.long)
        self._i = torch.tensor(self.dataset.items, dtype=torch.long)
        self._r = torch.tensor(self.dataset.ratings, dtype=torch.float32)   # read.py:124

    def __iter__(self):
        import torch
        n = len(self.dataset)
        if self.perms is None:
            order = torch.arange(n)
        else:
            order = torch.as_tensor(self.perms[self.calls], dtype=torch.long)
            self.calls += 1
        for s in range(0, n, self.batch):
            idx = order[s:s + self.batch]
            yield self._u[idx], self._i[idx], self._r[idx]


def make_model(n_user, n_item, k, P0, Q0):
    """Reference MF (utils.py:30-43) with injected initial weights."""
    import torch
    ns = load()
    m = ns.utils.MF(n_user, n_item, k)
    with torch.no_grad():
        m.user_mat.weight.copy_(torch.as_tensor(P0))
        m.item_mat.weight.copy_(torch.as_tensor(Q0))
    return m


def train_injected(arr3, n_user, n_item, k, P0, Q0, perms, batch, epochs,
                   lr=1e-3, lam=0.1, momentum=0.9, lr_decay=0.95):
    """Drive the reference's own ``baseTrain`` (utils.py:46-111) with the optimiser /
    scheduler construction of scratch.py:65-69,78-80.  Returns (model, losses)."""
    import torch
    from torch import nn, optim
    ns = load()
    model = make_model(n_user, n_item, k, P0, Q0)
    opt = optim.SGD(model.parameters(), lr=lr, weight_decay=lam, momentum=momentum)
    sched = optim.lr_scheduler.StepLR(opt, step_size=50, gamma=lr_decay)
    loader = PermLoader(arr3, batch, perms)
    loss_fn = nn.MSELoss(reduction="sum")
    losses = []
    for _ in range(epochs):
        loss, _ = ns.utils.baseTrain(loader, model, loss_fn, True, opt, "cpu", 0, "nor", [])
        sched.step()
        losses.append(float(loss))
    return model, losses


def base_test(arr3, models, batch):
    """Reference ``baseTest`` (utils.py:115-187) -> (rmse, ndcg, hr)."""
    from torch import nn
    ns = load()
    loader = PermLoader(arr3, batch, None)
    rmse, ndcg, hr = ns.utils.baseTest(loader, models, nn.MSELoss(reduction="sum"), "cpu", 0)
    return float(rmse), float(ndcg), float(hr)

"""One process per GPU over torch.distributed (NCCL on the box, gloo in CPU tests).

The reference has no multi-GPU code at all; this is the new shard->GPU
placement of SURVEY.md §8(e):
  * shard MF training / retraining: shard s lives on rank s mod world, NO communication;
  * Sinkhorn grouping: users row-sharded, all-reduce of the k column marginals per
    iteration and of the [k,d] centroid sums per outer iteration;
  * ensemble evaluation: every rank scores all test interactions against ITS shards' item
    tables, all-reduce(sum) of the partial score vector;
  * merge: owner rows are disjoint, all-reduce(sum) of zero-filled contributions.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch


class Dist:
    def __init__(self, group=None):
        import torch.distributed as td
        self.td = td
        self.group = group
        self.active = td.is_available() and td.is_initialized()
        self.world = td.get_world_size(group) if self.active else 1
        self.rank = td.get_rank(group) if self.active else 0

    # ---- placement
    def owner_of_shard(self, s: int) -> int:
        return s % self.world

    def my_shards(self, ids) -> List[int]:
        return [s for s in ids if self.owner_of_shard(s) == self.rank]

    def row_block(self, n: int):
        """Contiguous row range of this rank for row-sharded inputs."""
        per = -(-n // self.world)
        lo = min(n, per * self.rank)
        return lo, min(n, lo + per)

    # ---- collectives (no-ops at world == 1)
    def all_reduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        if self.world > 1:
            self.td.all_reduce(t, op={"sum": self.td.ReduceOp.SUM, "max": self.td.ReduceOp.MAX}[op],
                               group=self.group)
        return t

    def all_gather_into(self, out: torch.Tensor, part: torch.Tensor) -> torch.Tensor:
        """out [world * m, ...] <- every rank's `part` [m, ...] in rank order."""
        if self.world == 1:
            out.copy_(part)
        else:
            self.td.all_gather_into_tensor(out, part, group=self.group)
        return out

    def sum_int(self, v: int) -> int:
        if self.world == 1:
            return int(v)
        dev = "cuda" if self.td.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
        self.td.all_reduce(t, group=self.group)
        return int(t.item())

    def max_float(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        dev = "cuda" if self.td.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def broadcast_object(self, obj, src: int = 0):
        """`obj` of rank `src` on every rank (pickled; host-side bookkeeping such as the user grouping)."""
        if self.world == 1:
            return obj
        box = [obj if self.rank == src else None]
        self.td.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def barrier(self):
        if self.world > 1:
            self.td.barrier(group=self.group)


_default: Optional[Dist] = None


def init_from_env(backend: Optional[str] = None) -> Dist:
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun) if WORLD_SIZE > 1."""
    global _default
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend)
    _default = Dist()
    return _default


def get() -> Dist:
    global _default
    if _default is None or (_default.world == 1 and _default.td.is_initialized()):
        _default = Dist()
    return _default

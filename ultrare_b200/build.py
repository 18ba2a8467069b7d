"""Build libultrare_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m ultrare_b200.build [--force]

The shared library lands next to this file so that it travels with the repo
snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libultrare_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha1()
    h.update(" ".join(FLAGS).encode())
    paths = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    paths.append(os.path.join(os.path.dirname(HERE), "include", "ultrare_b200.h"))
    for p in paths:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    if os.path.exists(stamp):
        os.unlink(stamp)
    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, rc, out in ex.map(_compile, sources()):
            logs.append(f"== {os.path.basename(src)}\n{out}")
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{out}")
            objs.append(obj)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    if os.path.exists(LIB):
        os.unlink(LIB)              # never leave a stale library behind a failed link
    r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

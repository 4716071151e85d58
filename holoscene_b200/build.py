"""In-tree build of libhsb200.so (hand-written sm_100a CUDA, C ABI in include/hsb200.h).

    python -m holoscene_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
SO = os.path.join(PKG, "libhsb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, hdrs, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if force or _stale(obj, [src] + hdrs):
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(PKG), "include", "hsb200.h")]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdrs, force), srcs))
    if force or _stale(SO, objs):
        # link next to the target and rename: a reader (or a snapshot of the tree) never sees a half-written library
        subprocess.check_call([NVCC, "-shared", "-o", SO + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
        os.replace(SO + ".tmp", SO)
    if verbose:
        for o in objs:
            print(open(o + ".log").read())
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

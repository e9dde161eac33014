"""Compile the CUDA library in-tree for sm_100a (B200).  No torch headers: plain nvcc, one object per source file
(compiled in parallel, rebuilt only when stale), then one link.

    python simplex-gp_b200/csrc/build.py [--force] [--verbose]

The resulting ``simplex-gp_b200/libsgp_lattice.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
NAMES = ["sgp_lattice", "sgp_tiles", "sgp_grad", "sgp_groups", "sgp_solver", "sgp_ring", "sgp_filter"]
SRC = [os.path.join(HERE, n + ".cu") for n in NAMES]
HDR = [os.path.join(ROOT, "include", "sgp_lattice.h"), os.path.join(HERE, "sgp_common.cuh")]
OBJ_DIR = os.path.join(HERE, "build")
OUT = os.path.join(PKG, "libsgp_lattice.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # structure/value parity: no FMA contraction anywhere
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Wno-deprecated-declarations",
    "-I", os.path.join(ROOT, "include"),
]


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in deps)


def stale() -> bool:
    return _stale(OUT, SRC + HDR + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or stale()):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in SRC if force or _stale(_obj(s), [s] + HDR + [os.path.abspath(__file__)])]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(src), src]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        list(pool.map(compile_one, todo))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [_obj(s) for s in SRC]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Compile the CUDA library in-tree for sm_100a (B200).  No torch headers: plain nvcc, a few seconds.

    python simplex-gp_b200/csrc/build.py [--force] [--verbose]

The resulting ``simplex-gp_b200/libsgp_lattice.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
SRC = [os.path.join(HERE, "sgp_lattice.cu"), os.path.join(HERE, "sgp_tiles.cu"),
       os.path.join(HERE, "sgp_grad.cu"), os.path.join(HERE, "sgp_groups.cu"),
       os.path.join(HERE, "sgp_solver.cu")]
HDR = [os.path.join(ROOT, "include", "sgp_lattice.h"), os.path.join(HERE, "sgp_common.cuh")]
OUT = os.path.join(PKG, "libsgp_lattice.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # structure/value parity: no FMA contraction anywhere
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in SRC + HDR + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

#!/usr/bin/env python
"""Regenerate profiles/r2_sass_excerpt.txt from the built library (cuobjdump -sass; no GPU needed).

    python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simplex-gp_b200", "libsgp_lattice.so")
KERNELS = [   # (mangled-name regex, title)
    (r"_Z21sgp_splat_ring_kernelILi4ELb0ELb0EE", "sgp_splat_ring_kernel<VEC=4, !RAGGED, !SCAN>  (production splat for dense lattices, L = 8..64: TMA ring + reductions)"),
    (r"_Z21sgp_splat_rows_kernelILi4ELi8ELb0EE", "sgp_splat_rows_kernel<VEC=4, SEG=8, !RAGGED>  (one-shot splat: sparse / small lattices, L <= 4)"),
    (r"_Z21sgp_blur_group_kernelILi4ELi1ELi4ELi256ELb1ELi512EE", "sgp_blur_group_kernel<VEC=4, R=1, CHUNKS=4, 256 threads, FAST, 512 rows>  (production blur, L = 16, order 1)"),
    (r"_Z21sgp_slice_ring_kernelILi4ELb1ELb0EE", "sgp_slice_ring_kernel<VEC=4, FAST, !RAGGED>  (production slice, L >= 12)"),
    (r"_Z21sgp_splat_ring_kernelILi4ELb0ELb1EE", "sgp_splat_ring_kernel<VEC=4, !RAGGED, SCAN>  (optional: tile scan + stores, SGP_SPLAT_SCAN=1)"),
    (r"_Z20sgp_cg_update_kernelILi4EE", "sgp_cg_update_kernel<4>  (CG sweep: X += alpha P, R -= alpha AP, |R|^2 per column; 16-byte vectors)"),
]
SHOW = re.compile(r"UBLKCP|UBLKPF|SYNCS\.|ACQBULK|PREEXIT|LDGSTS|RED\.|UTMA")
COUNT = re.compile(r"^(LDG|STG|LDS|STS|RED|ATOM|SHFL|UBLK|SYNCS|LDGSTS|BAR|VOTE|FFMA|FADD|FMUL|ELECT|ACQBULK|PREEXIT|WARPSYNC|CCTL|DEPBAR|LDGDEPBAR|MATCH)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    print("# SASS of the hot kernels of libsgp_lattice.so (cuobjdump -sass, sm_100a, nvcc 12.9; round 2; profiles/sass_excerpt.py).")
    print("# Per kernel: instruction count, mnemonic histogram of the memory / synchronisation / async-copy instructions, and the")
    print("# lines that show the Blackwell/Hopper-era mechanisms (UBLKCP = cp.async.bulk (TMA), SYNCS.* = mbarrier, ACQBULK /")
    print("# PREEXIT = programmatic dependent launch, LDGSTS = cp.async, RED = red.global.add); at most 16 such lines per kernel.")
    for pat, title in KERNELS:
        hit = [b for b in blocks if re.match(pat, b)]
        if not hit:
            print(f"\n== {title}\n   (not found: {pat})")
            continue
        b = hit[0]
        name = b.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", b)
        hist = collections.Counter()
        for _, text in ins:
            op = re.sub(r"^@!?U?P\d+\s+", "", text.strip()).split()[0]
            if COUNT.match(op):
                hist[op] += 1
        print(f"\n== {title}\n   {name}\n   {len(ins)} instructions")
        print("   " + ", ".join(f"{k} x{v}" for k, v in hist.most_common(18)))
        shown = 0
        for addr, text in ins:
            if SHOW.search(text) and shown < 16:
                print(f"      /*{addr}*/  {text.strip()} ;")
                shown += 1


if __name__ == "__main__":
    sys.exit(main())

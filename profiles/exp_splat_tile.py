#!/usr/bin/env python
"""Round-2 experiment: tile size / ring depth / occupancy of the TMA-ring splat at the metric shape (config A).

    python profiles/exp_splat_tile.py

A persistent warp takes the tiles w, w + W, ...; with 256-entry tiles that is 7.4 tiles per warp, so the last round runs
with 42 % of the warps.  Smaller tiles even that out at the price of more bulk copies and mbarrier waits."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _ptr, _stream_ptr  # noqa: E402
import bench  # noqa: E402
from exp_ring import timed  # noqa: E402


def main():
    N, d, L = 1_000_000, 8, 16
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, d, generator=g).to(dev)
    lat = sg.Lattice(x, bench.COEFFS[("rbf", 1)])
    lib, st, rows, M = _capi.lib(), _stream_ptr(dev), lat.rows, lat.M
    Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
    buf = torch.zeros(M, L, device=dev)

    def splat(i):
        V = Vs[i % 4]
        _capi.check(lib.sgp_mvm_stage_splat_prezeroed(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V),
                                                      V.stride(0), L, _ptr(buf), L, st))
    ref = None
    for tile in (64, 128, 192, 256, 384, 512):
        for stages in (2, 3, 4):
            for occ in (0, 3):
                os.environ.update(SGP_SPLAT_TILE=str(tile), SGP_SPLAT_STAGES=str(stages), SGP_RING_OCC=str(occ))
                try:
                    buf.zero_()
                    splat(0)
                    torch.cuda.synchronize()
                    if ref is None:
                        ref = buf.clone()
                    rel = float((buf - ref).abs().max() / ref.abs().max())
                    us = timed(splat, 30)
                    print(f"tile {tile:4d} stages {stages} occ {occ}: {us:6.1f} us  rel {rel:.1e}", flush=True)
                except Exception as exc:
                    print(f"tile {tile} stages {stages} occ {occ}: {str(exc)[:100]}", flush=True)


if __name__ == "__main__":
    main()

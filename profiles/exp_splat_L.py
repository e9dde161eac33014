#!/usr/bin/env python
"""Round-2 experiment: one-shot vs TMA-ring row-sorted splat over the number of right-hand sides at the metric lattice
(N = 1M, d = 8).  Decides the `ring_pays` rule in csrc/sgp_tiles.cu::splat_rows_impl.

    python profiles/exp_splat_L.py [--Ls 1,4,8,12,16,32]"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--Ls", default="1,4,8,12,16,32")
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--d", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(args.N, args.d, generator=g).to(dev)
    lat = sg.Lattice(x, bench.COEFFS[("rbf", 1)])
    lib, st, rows, M, N = _capi.lib(), _stream_ptr(dev), lat.rows, lat.M, args.N
    for L in [int(s) for s in args.Ls.split(",")]:
        Lv = (L + 3) // 4 * 4 if L > 4 else L
        Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
        buf = torch.empty(M, Lv, device=dev)

        def splat(i):
            V = Vs[i % 4]
            _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V), V.stride(0),
                                           L, _ptr(buf), Lv, st))
        res = {}
        for ring in (0, 1):
            os.environ["SGP_RING_SPLAT"] = str(ring)
            os.environ["SGP_RING_FORCE"] = str(ring)
            res[ring] = timed(splat, 30)
            ref = buf.clone() if ring == 0 else ref
        os.environ.pop("SGP_RING_FORCE")
        rel = float((buf - ref).abs().max() / ref.abs().max())
        print(f"L={L} Lv={Lv}: one-shot {res[0]:.1f} us, ring {res[1]:.1f} us, rel {rel:.1e}", flush=True)


if __name__ == "__main__":
    main()

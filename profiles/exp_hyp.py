#!/usr/bin/env python
"""Where does the one-shot row-sorted splat spend its time?  Times it with pieces switched off (SGP_SPLAT_DBG)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _ptr, _stream_ptr  # noqa: E402
import bench  # noqa: E402
from profiles.exp_ring import timed  # noqa: E402

w = bench.WORKLOADS["A"]
N, d, L = w["N"], w["d"], w["L"]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
x = torch.randn(N, d, generator=g).to(dev)
lat = sg.Lattice(x, bench.COEFFS[("rbf", 1)])
M, rows = lat.M, lat.rows
lib, st = _capi.lib(), _stream_ptr(dev)
Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
buf0 = torch.empty(M, L, device=dev)


def splat(i):
    V = Vs[i % 4]
    _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V), V.stride(0), L,
                                   _ptr(buf0), L, st))


os.environ["SGP_RING_SPLAT"] = "0"
for name, dbg in [("baseline", 0), ("no reductions", 1), ("no memset", 2), ("no reductions, no memset", 3),
                  ("V folded to 64 KB", 10 << 8), ("V folded to 1 MB", 14 << 8), ("V folded to 16 MB", 18 << 8),
                  ("V folded to 1 MB, no reductions, no memset", (14 << 8) | 3),
                  ("V folded to 64 KB, no reductions, no memset", (10 << 8) | 3)]:
    os.environ["SGP_SPLAT_DBG"] = str(dbg)
    print(json.dumps({"variant": name, "splat_us": round(timed(splat, 30), 2)}), flush=True)
for seg in (4, 16):
    pass

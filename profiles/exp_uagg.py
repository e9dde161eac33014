#!/usr/bin/env python
"""One-shot row-sorted splat with / without the warp-uniform aggregation of long rows (SGP_SPLAT_UAGG)."""
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

for wl in sys.argv[1:] or ["A"]:
    w = bench.WORKLOADS[wl]
    N, d, L = w["N"], w["d"], w["L"]
    dev = torch.device("cuda", 0)
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
    lat = sg.Lattice(x, bench.COEFFS[(w["kernel"], w["order"])])
    M, rows = lat.M, lat.rows
    lib, st = _capi.lib(), _stream_ptr(dev)
    Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
    outs = [torch.empty(N, L, device=dev) for _ in range(4)]
    buf0 = torch.empty(M, L, device=dev)

    def splat(i):
        V = Vs[i % 4]
        _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V), V.stride(0), L,
                                       _ptr(buf0), L, st))

    res = {}
    for u in (0, 1):
        os.environ["SGP_SPLAT_UAGG"] = str(u)
        splat(0)
        res[u] = buf0.clone()
        t = timed(splat, 30)
        graphs = [lat.capture(Vs[k], outs[k]) for k in range(4)]
        g = timed(lambda i: graphs[i % 4].replay(), 300, warm=20)
        del graphs
        print(json.dumps({"workload": wl, "uagg": u, "splat_us": round(t, 2), "graph_mvm_us": round(g, 2)}), flush=True)
    print(json.dumps({"workload": wl, "rel_diff": float((res[1] - res[0]).norm() / res[0].norm())}), flush=True)
    del lat, Vs, outs, buf0, res

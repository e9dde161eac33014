#!/usr/bin/env python
"""Blur-group geometry sweep at the metric shape: rows per CTA x axes per group -> MVM time (CUDA-graph replay)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
import bench  # noqa: E402
from profiles.exp_ring import timed  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "A"]
N, d, L = w["N"], w["d"], w["L"]
dev = torch.device("cuda", 0)
x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
outs = [torch.empty(N, L, device=dev) for _ in range(4)]
for rows in (256, 512, 1024):
    for axes in (None, 2, 3, 4, 5):
        try:
            lat = sg.Lattice(x, bench.COEFFS[(w["kernel"], w["order"])], group_rows=rows, group_axes=axes)
            if lat.groups is None:
                print(json.dumps({"rows": rows, "axes": axes, "groups": None}), flush=True)
                continue
            graphs = [lat.capture(Vs[k], outs[k]) for k in range(4)]
            t = timed(lambda i: graphs[i % 4].replay(), 300, warm=20)
            print(json.dumps({"rows": rows, "axes": axes, "ranges": [(g["j0"], g["j1"], g["max_class"], g["n_batches"]) for g in lat.groups["list"]],
                              "mvm_us": round(t, 2)}), flush=True)
            del graphs, lat
        except Exception as exc:
            print(json.dumps({"rows": rows, "axes": axes, "error": str(exc)[:200]}), flush=True)

#!/usr/bin/env python
"""Lattice build time at a BASELINE shape, first build (safe hash-table size) vs later builds (table sized from the last M)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
import bench  # noqa: E402

for wl in sys.argv[1:] or ["A"]:
    w = bench.WORKLOADS[wl]
    N, d = w["N"], w["d"]
    dev = torch.device("cuda", 0)
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
    c = bench.COEFFS[(w["kernel"], w["order"])]
    times, caps = [], []
    for i in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = sg.Lattice(x, c)
        e1.record()
        torch.cuda.synchronize()
        times.append(round(e0.elapsed_time(e1), 3))
        caps.append(lat.hash_capacity)
        del lat
    bare = []
    for i in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = sg.Lattice(x, c, build_groups=False, build_rows=False)
        e1.record()
        torch.cuda.synchronize()
        bare.append(round(e0.elapsed_time(e1), 3))
        del lat
    print(json.dumps({"workload": wl, "build_ms": times, "hash_capacity": caps, "without_groups_and_rows_ms": bare}), flush=True)

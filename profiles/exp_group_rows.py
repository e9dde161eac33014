#!/usr/bin/env python
"""Round-2 experiment: rows per blur-group CTA (``Lattice(group_rows=...)``) against the MVM time (CUDA-graph replay) at
configs B and A.  Asked because 616 batches of 512 rows fill the 444 CTA slots of a B200 1.39 times (config B); result:
the larger batch wins anyway (B: 76.0 us at 512, 77.3 at 448, 83.9 at 384, 91.3 at 256; A: 170.7 / 180.5 / 184.4 / 193.4) --
the per-CTA cost (two load round trips, two barriers per pass) outweighs the wave quantisation.

    python profiles/exp_group_rows.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simplex_gp_b200 as sg
import bench
from exp_ring import timed
for wl in ("B", "A"):
    w = bench.WORKLOADS[wl]
    N, d, L = w["N"], w["d"], w["L"]
    dev = torch.device("cuda", 0)
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
    Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
    outs = [torch.empty(N, L, device=dev) for _ in range(4)]
    for rows in (192, 256, 320, 384, 448, 512):
        lat = sg.Lattice(x, bench.COEFFS[(w["kernel"], w["order"])], group_rows=rows)
        graphs = [lat.capture(Vs[k], outs[k]) for k in range(4)]
        t = timed(lambda i: graphs[i % 4].replay(), 300, warm=20)
        print(json.dumps({"workload": wl, "rows": rows, "ranges": [(g["j0"], g["j1"], g["max_class"], g["n_batches"]) for g in lat.groups["list"]], "mvm_us": round(t, 2)}), flush=True)
        del graphs, lat

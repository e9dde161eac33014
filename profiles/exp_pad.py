#!/usr/bin/env python
"""Ring slice: dense replay table vs the padded table read in pairs (SGP_REPLAY_PAD)."""
import ctypes as C
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
    Vs = [torch.randn(N, L, device=dev) for _ in range(4)]
    outs = [torch.empty(N, L, device=dev) for _ in range(4)]
    ref = None
    for pad in (0, 1):
        os.environ["SGP_REPLAY_PAD"] = str(pad)
        lat = sg.Lattice(x, bench.COEFFS[(w["kernel"], w["order"])])
        lib, st = _capi.lib(), _stream_ptr(dev)
        Lv = (L + 3) // 4 * 4 if L > 4 else L
        buf = torch.randn(lat.M, Lv, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
        v_out = lat._view(lat._table(False, True), None, lat.exact)

        def slice_(i):
            o = outs[i % 4]
            _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(buf), Lv, _ptr(o), o.stride(0), L, st))

        slice_(0)
        got = outs[0].clone()
        if ref is None:
            ref = got
        t = timed(slice_, 30)
        graphs = [lat.capture(Vs[k], outs[k]) for k in range(4)]
        g = timed(lambda i: graphs[i % 4].replay(), 300, warm=20)
        print(json.dumps({"workload": wl, "pad": pad, "stride": int(lat._table(False, True).shape[1]), "slice_us": round(t, 2),
                          "graph_mvm_us": round(g, 2), "same_bits": bool(torch.equal(got, ref))}), flush=True)
        del graphs, lat

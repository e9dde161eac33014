#!/usr/bin/env python
"""One-shot slice at narrow rows: dense [N, d+1] replay table vs the transposed [d+1, N] one."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _ptr, _stream_ptr  # noqa: E402
from profiles.exp_ring import timed  # noqa: E402

N, d = 1_000_000, int(os.environ.get("SGP_D", 8))
dev = torch.device("cuda", 0)
x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
lat = sg.Lattice(x, [0.34608543, 1.0, 0.34608543])
M = lat.M
lib, st = _capi.lib(), _stream_ptr(dev)
views = {"dense": lat._view(lat._table(False, True), None, lat.exact),
         "transposed": lat._view(lat._table(False, True, True), None, lat.exact, transposed=True)}
for L in (1, 2, 4, 8):
    buf = torch.randn(M, L, device=dev)
    outs = [torch.empty(N, L, device=dev) for _ in range(4)]
    rec, ref = {"L": L}, None
    for name, v in views.items():
        def slice_(i):
            o = outs[i % 4]
            _capi.check(lib.sgp_slice(C.byref(v), _ptr(buf), L, _ptr(o), o.stride(0), L, st))
        slice_(0)
        got = outs[0].clone()
        ref = got if ref is None else ref
        rec[name] = round(timed(slice_, 30), 1)
        rec[name + "_same"] = bool(torch.equal(got, ref))
    print(json.dumps(rec), flush=True)

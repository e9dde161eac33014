"""Minimal driver for ncu captures of the stage kernels at the metric shape (one launch of each per repetition).

    ncu --set full --clock-control none --import-source on -k regex:'sgp_(splat|slice)_' -s 4 -c 4 \
        -o gpurun_out/ring python profiles/ncu_ring.py
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _ptr, _stream_ptr  # noqa: E402

N, d, L = int(os.environ.get("SGP_N", 1_000_000)), int(os.environ.get("SGP_D", 8)), int(os.environ.get("SGP_L", 16))
torch.manual_seed(0)
dev = torch.device("cuda", 0)
x = torch.randn(N, d, device=dev)
lat = sg.Lattice(x, [0.34608543, 1.0, 0.34608543])
lib, st = _capi.lib(), _stream_ptr(dev)
M, rows = lat.M, lat.rows
Vs = [torch.randn(N, L, device=dev) for _ in range(2)]
outs = [torch.empty(N, L, device=dev) for _ in range(2)]
buf0, buf1 = torch.empty(M, L, device=dev), torch.randn(M, L, device=dev)
v_out = lat._view(lat._table(False, True), None, lat.exact)
for i in range(int(os.environ.get("SGP_REPS", 4))):
    V, o = Vs[i % 2], outs[i % 2]
    _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V), V.stride(0), L,
                                   _ptr(buf0), L, st))
    _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(buf1), L, _ptr(o), o.stride(0), L, st))
torch.cuda.synchronize()
print("M", M, "checksum", float(outs[0].double().sum()), float(buf0.double().sum()))

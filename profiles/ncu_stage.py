"""Driver for ncu captures of single stages in several configurations (config A lattice).

    ncu --set full --clock-control none --import-source on -k regex:sgp_slice -c 4 -o gpurun_out/x python profiles/ncu_stage.py slice
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "slice"
N, d, L = 1_000_000, 8, 16
torch.manual_seed(0)
x = torch.randn(N, d, device="cuda")
v = torch.randn(N, L, device="cuda")
lat = sg.Lattice(x, [0.34608543, 1.0, 0.34608543], sort_points=True)
lib = _capi.lib()
buf0, buf1 = lat._scratch(L)
st = _stream_ptr(lat.device)
out = torch.empty(N, L, device="cuda")
lat.mvm(v, out=out)
torch.cuda.synchronize()
for tr in (False, True):
    for ex in (True, False):
        view = lat._view(lat._table(False, True), None, ex, True) if tr else lat._view(exact=ex)
        if what == "slice":
            _capi.check(lib.sgp_slice(C.byref(view), _ptr(buf1), L, _ptr(out), out.stride(0), st))
        else:
            _capi.check(lib.sgp_splat(C.byref(view), _ptr(v), v.stride(0), L, _ptr(buf0), 1, st))
torch.cuda.synchronize()
print("done", what)

"""Minimal driver for ncu captures: build the config-A lattice, run a few MVMs.

    ncu --set full --clock-control none --import-source on -k regex:"sgp_(splat|blur|slice)" -s 10 -c 5 -o gpurun_out/mvm python profiles/ncu_mvm.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402

N, d, L = int(os.environ.get("SGP_N", 1_000_000)), int(os.environ.get("SGP_D", 8)), int(os.environ.get("SGP_L", 16))
mode = int(os.environ.get("SGP_SPLAT", 0))   # 0 = the production chain
torch.manual_seed(0)
x = torch.randn(N, d, device="cuda")
v = torch.randn(N, L, device="cuda")
blur = os.environ.get("SGP_BLUR", "auto")
lat = sg.Lattice(x, [0.34608543, 1.0, 0.34608543], build_csr=(mode == 2))
out = torch.empty(N, L, device="cuda")
for _ in range(int(os.environ.get("SGP_REPS", 3))):
    lat.mvm(v, out=out, mode=mode, blur=blur)
torch.cuda.synchronize()
print("M", lat.M, "checksum", float(out.double().sum()))

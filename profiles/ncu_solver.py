"""Driver for an ncu capture of the CG sweeps (csrc/sgp_solver.cu) at N = 1M, 12 columns:

    ncu --set full --clock-control none -k regex:sgp_cg_ -s 10 -c 5 -o gpurun_out/solver python profiles/ncu_solver.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import gp  # noqa: E402

torch.manual_seed(0)
N, d = 1_000_000, 8
x = torch.randn(N, d, device="cuda")
B = torch.randn(N, 11, device="cuda")
op = sg.RBFLattice(ard_num_dims=d, order=1).cuda()(x)
s, noise = torch.tensor(0.7, device="cuda"), torch.tensor(0.3, device="cuda")
with torch.no_grad():
    X, al, be = gp.batched_cg(lambda V: s * op.matmul(V) + noise * V, B, tol=1e-3, max_iter=6, matmul=op.matmul, scale=s,
                              shift=noise)
torch.cuda.synchronize()
print("iterations", al.shape[0], "checksum", float(X.double().sum()))

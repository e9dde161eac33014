#!/usr/bin/env python
"""Round-2 experiment: device time of one CG iteration on the metric lattice (N = 1M, d = 8, 12-column block), with the
sweep after the product folded into the slice (default) or as its own launch (SGP_CG_FUSE=0).

    python profiles/exp_cg_iteration.py [iterations=40]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import gp  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N, d, L = 1_000_000, 8, 11
torch.manual_seed(0)
x = torch.randn(N, d, device="cuda")
kern = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
op = kern(x)
B = torch.randn(N, L, device="cuda")
s_, n_ = torch.tensor(1.0, device="cuda"), torch.tensor(0.1, device="cuda")
for fuse in ("0", "1", "0", "1"):
    os.environ["SGP_CG_FUSE"] = fuse
    with torch.no_grad():
        gp.batched_cg(None, B, tol=0.0, max_iter=5, matmul=op.matmul, scale=s_, shift=n_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        X, al, be = gp.batched_cg(None, B, tol=0.0, max_iter=iters, matmul=op.matmul, scale=s_, shift=n_)
        e1.record()
        torch.cuda.synchronize()
    print(f"fuse={fuse}: {e0.elapsed_time(e1) / al.shape[0] * 1e3:7.1f} us per CG iteration over {al.shape[0]} iterations "
          f"(checksum {float(X.double().abs().sum()):.6e})", flush=True)

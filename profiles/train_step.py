"""One marginal-likelihood training step (CG solve on [y | 10 probes] + Lanczos log-det + lengthscale gradient), the
caller either side of the MVM (BASELINE.json configs[2]: elevators-shaped N=16.6k, d=18, RBFLattice order 1, 11 RHS),
timed on the GPU with this package's solver (GPyTorch is not available).

    python profiles/train_step.py [N d [stop min_iter]]

stop / min_iter: the CG stopping rule (gp.batched_cg): "all" 0 = every column's relative residual under the tolerance
(the default of this package); "mean" 20 = GPyTorch's rule, i.e. what the reference's settings train with.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import gp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16_600
d = int(sys.argv[2]) if len(sys.argv) > 2 else 18
stop = sys.argv[3] if len(sys.argv) > 3 else "all"
min_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 0
torch.manual_seed(0)
x = torch.randn(N, d, device="cuda")
y = torch.tanh(x[:, 0]) + 0.5 * torch.sin(x[:, 1:3].sum(1)) + 0.1 * torch.randn(N, device="cuda")
kernel = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
model = gp.ExactGPModel(x, y, kernel, max_cholesky_size=0).cuda()
opt = torch.optim.Adam(model.parameters(), lr=0.1)
probes = torch.randn(N, 10, device="cuda").sign()
times, values, iters = [], [], []
for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opt.zero_grad()
    stats = {}
    value, surrogate = model.mll(probes=probes, tol=1.0 if it else 1e-2, max_iter=100, stats=stats, stop=stop,
                                 min_iter=min_iter)   # reference: cg_tolerance 1.0
    (-surrogate).backward()
    opt.step()
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    values.append(value)
    iters.append(stats.get("cg_iterations"))
print(f"N={N} d={d} stop={stop} min_iter={min_iter}: CG iterations {iters}; MLL per datum {values[0]:.4f} -> {values[-1]:.4f}; step time ms {[round(t * 1e3, 1) for t in times]}; "
      f"lattice builds {sg.lattice_cache.builds}, cache hits {sg.lattice_cache.hits}")

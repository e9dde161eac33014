"""Where one N = 1M training step goes: kernel-time totals (torch.profiler / CUPTI) against the wall time of the step.

    python profiles/train_step_breakdown.py [N d]
"""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import gp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
x = torch.randn(N, d, device="cuda")
y = torch.tanh(x[:, 0]) + 0.5 * torch.sin(x[:, 1:3].sum(1)) + 0.1 * torch.randn(N, device="cuda")
kernel = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
model = gp.ExactGPModel(x, y, kernel, max_cholesky_size=0).cuda()
opt = torch.optim.Adam(model.parameters(), lr=0.1)
probes = torch.randn(N, 10, device="cuda").sign()


def step(it):
    opt.zero_grad()
    value, surrogate = model.mll(probes=probes, tol=1.0 if it else 1e-2, max_iter=100)
    (-surrogate).backward()
    opt.step()


for it in range(5):
    step(it)
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(5, 10):
    step(it)
torch.cuda.synchronize()
lats = [e for e in sg.lattice_cache._entries.values()]
Ms = [getattr(e, "M", None) or getattr(e[0] if isinstance(e, (tuple, list)) else e, "M", None) for e in lats]
print(f"step wall ms {(time.perf_counter() - t0) / 5 * 1e3:.2f}; cached lattice rows M = {Ms}; "
      f"lengthscale {kernel.lengthscale.detach().flatten()[:3].tolist()}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for it in range(10, 13):
        step(it)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[2])
total = sum(r[2] for r in rows)
print(f"GPU kernel time per step {total / 3 / 1e3:.2f} ms over {sum(r[1] for r in rows) / 3:.0f} launches")
for k, c, t in rows[:40]:
    print(f"{t / 3:9.1f} us {c / 3:7.1f} x  {k[:110]}")

"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck), one tool per run:

    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402

torch.manual_seed(0)
for (N, d, L, c) in ((3000, 8, 16, [0.34608543, 1.0, 0.34608543]), (1500, 5, 3, [0.1, 0.5, 1.0, 0.5, 0.1]),
                     (800, 18, 11, [0.34608543, 1.0, 0.34608543]), (64, 1, 2, [0.5, 1.0, 0.5])):
    x = torch.randn(N, d, device="cuda")
    v = torch.randn(N, L, device="cuda")
    lat = sg.Lattice(x, c, build_csr=True, build_tiles=True, sort_points=True)
    ref = lat.mvm(v, mode=2, blur="axis", exact=True)
    for kw in (dict(), dict(mode=1), dict(mode=3), dict(mode=4, blur="axis"), dict(sorted=True, mode=1), dict(exact=True)):
        out = lat.mvm(v, **kw)
        err = float((out - ref).norm() / ref.norm())
        assert err < 1e-5, (N, d, kw, err)
    g = lat.capture(v, torch.empty_like(v))
    g.replay()
    # the TMA-ring splat / slice forced onto these small shapes (both reduction forms), ragged blocks through the
    # zero-padded copy, eager and captured; then the one-call C entries (device and host pointers)
    os.environ.update(SGP_RING_FORCE="1", SGP_PAD_SRC="1")
    for scan in ("0", "1"):
        os.environ["SGP_SPLAT_SCAN"] = scan
        out = lat.mvm(v)
        err = float((out - ref).norm() / ref.norm())
        assert err < 1e-5, (N, d, "ring", scan, err)
        o2 = torch.empty_like(v)
        g2 = lat.capture(v, o2)
        g2.replay()
        g2.replay()
        torch.cuda.synchronize()
        assert float((o2 - ref).norm() / ref.norm()) < 1e-5, (N, d, "ring graph", scan)
    for k_ in ("SGP_RING_FORCE", "SGP_PAD_SRC", "SGP_SPLAT_SCAN"):
        os.environ.pop(k_)
    f1 = sg.filter(v, x, torch.tensor(c))
    f2 = sg.filter(v.cpu(), x.cpu(), torch.tensor(c))
    assert float((f1 - ref).norm() / ref.norm()) < 1e-5 and float((f2.cuda() - ref).norm() / ref.norm()) < 1e-5

    class KF:
        def get_coeffs(self):
            return torch.tensor(c)

        def get_deriv_coeffs(self):
            return torch.tensor(c)

    # lattice extension (sgp_hash_seed / sgp_hash_extend / sgp_*_extension) and the point-subset row tables (fillers)
    k = N // 2
    ext = sg.Lattice(x[:k].contiguous(), c).extend(x[k:].contiguous())
    assert torch.equal(ext.keys, lat.keys) and torch.equal(ext.replay, lat.replay)
    part = sg.Lattice.from_arrays(c, lat.replay[: k // 2].contiguous(), lat.keys, lat.nbr)
    part.mvm(v[: k // 2].contiguous())
    # CG sweeps (csrc/sgp_solver.cu) on the lattice operator
    from simplex_gp_b200 import gp
    kern = sg.RBFLattice(ard_num_dims=d, order=1).cuda() if len(c) == 3 else sg.MaternLattice(nu=1.5, order=2).cuda()
    op = kern(x)
    s_, n_ = torch.tensor(0.8, device="cuda"), torch.tensor(0.5, device="cuda")
    with torch.no_grad():
        X, al, be = gp.batched_cg(lambda V: s_ * op.matmul(V) + n_ * V, v, tol=1e-3, max_iter=50, matmul=op.matmul,
                                  scale=s_, shift=n_)
    assert torch.isfinite(X).all() and al.shape[0] >= 1

    xr = x.clone().requires_grad_(True)
    vr = v.clone().requires_grad_(True)
    sg.LatticeFilterGeneral.apply(vr, xr, KF()).sum().backward()
    assert torch.isfinite(xr.grad).all() and torch.isfinite(vr.grad).all()
torch.cuda.synchronize()
print("sanitize_small ok")

"""Cost of one prediction-shaped rectangular product K(test, train) V at the metric configuration (N = 1M training
points, d = 8, 16 columns) for several test-batch sizes: the reference's way (union lattice built from nothing,
bilateral_kernel.py:150-156) against extending the cached training lattice (Lattice.extend).

    gpurun -- python profiles/predict_step.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    torch.manual_seed(0)
    N, d, L = 1_000_000, 8, 16
    c = [0.34608543, 1.0, 0.34608543]
    x = torch.randn(N, d, device="cuda")
    v = torch.randn(N, L, device="cuda")
    base = sg.Lattice(x, c)
    print(f"training lattice: N={N} M={base.M}; build {timed(lambda: sg.Lattice(x, c)):.2f} ms")
    for n_test in (1_000, 10_000, 100_000, 1_000_000):
        xt = torch.randn(n_test, d, device="cuda")
        union = torch.cat([x, xt])
        t_full = timed(lambda: sg.Lattice(union, c))
        t_ext = timed(lambda: base.extend(xt))
        t_ext_plain = timed(lambda: base.extend(xt, build_groups=False, build_rows=False))
        lat = base.extend(xt)
        src = torch.cat([v, torch.zeros(n_test, L, device="cuda")])
        t_mvm = timed(lambda: lat.mvm(src), reps=20)
        plain = base.extend(xt, build_groups=False, build_rows=False)
        t_plain = timed(lambda: plain.mvm(src), reps=20)
        print(f"n_test={n_test:>8}: union build {t_full:6.2f} ms | extend {t_ext:6.2f} ms | extend without derived "
              f"tables {t_ext_plain:6.2f} ms | product {t_mvm:5.2f} ms, on the neighbour table alone {t_plain:5.2f} ms "
              f"| M_union={lat.M}")
    # the operator as a user calls it: K(test, train) @ V with the training lattice cached (every call divides x by the
    # lengthscale afresh, so the cache finds the lattices by value)
    k = sg.RBFLattice(ard_num_dims=d, order=1).cuda()

    def once(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    with torch.no_grad():
        k(x).matmul(v)
        k(torch.randn(100, d, device="cuda"), x).matmul(v)          # first use of the extension kernels (module load)
        for n_test in (10_000, 100_000):
            xt = torch.randn(n_test, d, device="cuda")
            first = once(lambda: k(xt, x).matmul(v))
            rest = [once(lambda: k(xt, x).matmul(v)) for _ in range(24)]
            print(f"operator, n_test={n_test}: first product {first:.2f} ms (extends the training lattice), then "
                  f"{sorted(rest)[len(rest) // 2]:.2f} ms median, {max(rest):.2f} ms max (the product that builds the "
                  f"postponed tables); cache: {sg.lattice_cache.builds} builds, {sg.lattice_cache.extensions} "
                  f"extensions, {sg.lattice_cache.content_hits} hits by value")


if __name__ == "__main__":
    main()

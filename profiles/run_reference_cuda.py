"""Optional second baseline: the reference's own CUDA extension
(gpytorch_lattice_kernel/cuda/permutohedral_cuda*.c*), compiled UNMODIFIED for sm_100 by the recipe below into the
git-ignored oracle/_ref/cuda_build/ (build container, no GPU needed), run here on the B200 beside this package.

    # build (container with /root/reference):   python profiles/run_reference_cuda.py --build
    # run (GPU box):                             python profiles/run_reference_cuda.py

It is a cross-check at allclose level only: the reference's CUDA path is not bit-compatible with its CPU path
(different rounding rule in the simplex search, double-promoted arithmetic).
"""
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BUILD = os.path.join(ROOT, "oracle", "_ref", "cuda_build")
NAME = "sgp_ref_gpu_lattice"


def build():
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    d = "/root/reference/gpytorch_lattice_kernel/cuda"
    os.makedirs(BUILD, exist_ok=True)
    load(name=NAME, sources=[os.path.join(d, "permutohedral_cuda.cpp"), os.path.join(d, "permutohedral_cuda_kernel.cu")],
         build_directory=BUILD, verbose=True, is_python_module=False)


def load_module():
    so = os.path.join(BUILD, NAME + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    if "--build" in sys.argv:
        return build()
    ref = load_module()
    if ref is None:
        print(json.dumps({"unavailable": "oracle/_ref/cuda_build/sgp_ref_gpu_lattice.so not built"}))
        return
    import simplex_gp_b200 as sg
    res = {}
    c = torch.tensor([0.34608543, 1.0, 0.34608543], device="cuda")
    for N, d, L in ((100_000, 8, 16), (1_000_000, 8, 1), (1_000_000, 8, 16)):
        g = torch.Generator().manual_seed(0)
        x = torch.randn(N, d, generator=g).cuda()
        v = torch.randn(N, L, generator=g).cuda()
        ours = sg.filter(v, x, c)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ours = sg.filter(v, x, c)
        torch.cuda.synchronize()
        t_ours = time.perf_counter() - t0
        theirs = ref.filter(v, x, c)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        theirs = ref.filter(v, x, c)
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
        rel = float((ours - theirs).norm() / theirs.norm())
        res[f"N={N},d={d},L={L}"] = {"reference_cuda_s": t_ref, "ours_filter_s": t_ours, "speedup": t_ref / t_ours,
                                     "rel_l2_difference": rel}
        print(f"N={N} d={d} L={L}: reference CUDA filter {t_ref * 1e3:.1f} ms, ours (build + MVM) {t_ours * 1e3:.2f} ms, "
              f"x{t_ref / t_ours:.0f}; rel L2 difference {rel:.2e}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate the committed golden vectors from the REFERENCE's own code (TEST INFRASTRUCTURE ONLY).

Run in the build container, where ``/root/reference`` exists:

    python tests/golden/make_golden.py

What runs the reference:
  * ``oracle/_ref/sgp_ref_lattice.so``  -- the unmodified reference header
    ``gpytorch_lattice_kernel/cpp/permutohedral.h`` compiled where it lies, plus
    ``oracle/ref_harness.cpp`` that exposes its intermediates (see oracle/build_oracle.py);
  * ``gpytorch_lattice_kernel/bilateral_kernel.py`` imported from ``/root/reference`` with
    ``gpytorch.kernels.Kernel`` / ``gpytorch.lazy.LazyTensor`` stubbed (GPyTorch is not installed
    here) and its native ``filter`` pointed at the module above instead of a second JIT build
    of the same source (``LatticeFilterGeneral.method``, bilateral_kernel.py:60).

Files written next to this script (small, committed):
  coeffs.json        stencil coefficients of DiscretizedKernelFN (bilateral_kernel.py:162-181)
  kat.npz            the two known-answer cases of SURVEY.md section 8c (d=2 toy, Snelson-1D)
  structure.npz      full lattice intermediates of a few seeded cases
  autograd.npz       forward / grad_source / grad_reference of LatticeFilterGeneral
  large.json         sha256 of every intermediate for cases with M > 16383, twice: of the reference with its
                     hash-growth statement re-ordered (see oracle/build_oracle.py) -- what the product
                     matches -- and of the UNMODIFIED reference -- what the oracle's reference-table mode
                     matches --, plus the deviation between the two
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_oracle  # noqa: E402

REF_ROOT = build_oracle.REF_ROOT

FIELDS = ("greedy", "rank", "offsets", "weights", "keys", "splatted", "blurred", "out", "scale")


def import_reference_python():
    gp, k, lz = (types.ModuleType(n) for n in ("gpytorch", "gpytorch.kernels", "gpytorch.lazy"))
    k.Kernel = type("Kernel", (nn.Module,), {})
    lz.LazyTensor = type("LazyTensor", (), {})
    sys.modules.update({"gpytorch": gp, "gpytorch.kernels": k, "gpytorch.lazy": lz})
    sys.path.insert(0, REF_ROOT)
    from gpytorch_lattice_kernel import bilateral_kernel as bk
    return bk


def inputs(N, d, L, seed, dist="randn", scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, d, generator=g) if dist == "randn" else torch.rand(N, d, generator=g)
    v = torch.randn(N, L, generator=g)
    return (x * scale).contiguous(), v.contiguous()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    build_oracle.build_ref()
    ref = build_oracle.load_ref(fixed=False)
    ref_fixed = build_oracle.load_ref(fixed=True)
    assert ref is not None and ref_fixed is not None, "reference build unavailable"
    bk = import_reference_python()
    bk.LatticeFilterGeneral.method = ref.filter

    # ---- 1. coefficients ------------------------------------------------------------------
    import contextlib
    import io
    coeffs = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for order in (1, 2, 3):
            dk = bk.DiscretizedKernelFN(bk.rbf, order)
            coeffs[f"rbf/{order}"] = {"forward": dk.get_coeffs().tolist(), "deriv": dk.get_deriv_coeffs().tolist()}
            for nu in (1.5, 2.5):
                dk = bk.DiscretizedKernelFN(lambda d2, nu=nu: bk.Matern.apply(d2, nu), order)
                coeffs[f"matern{nu}/{order}"] = {"forward": dk.get_coeffs().tolist(),
                                                 "deriv": dk.get_deriv_coeffs().tolist()}
    with open(os.path.join(HERE, "coeffs.json"), "w") as f:
        json.dump(coeffs, f, indent=1)
    rbf1 = torch.tensor(coeffs["rbf/1"]["forward"])
    rbf2 = torch.tensor(coeffs["rbf/2"]["forward"])
    mat15_2 = torch.tensor(coeffs["matern1.5/2"]["forward"])
    mat15_3 = torch.tensor(coeffs["matern1.5/3"]["forward"])

    # ---- 2. known-answer cases ------------------------------------------------------------
    kat = {}
    x = torch.tensor([[0.0, 0.0], [0.3, -1.2], [2.5, 0.7], [0.31, -1.19]])
    v = torch.tensor([[1.0], [2.0], [-1.0], [0.5]])
    c = torch.tensor([0.5, 1.0, 0.5])
    for name, val in zip(FIELDS, ref.structure(v, x, c)):
        kat[f"toy/{name}"] = val.numpy()
    kat["toy/x"], kat["toy/v"], kat["toy/coeffs"] = x.numpy(), v.numpy(), c.numpy()
    sn = np.loadtxt(os.path.join(REF_ROOT, "notebooks", "snelson.csv"), delimiter=",", skiprows=1).astype(np.float32)
    x = torch.from_numpy(np.ascontiguousarray(sn[:, :1]))
    v = torch.from_numpy(np.ascontiguousarray(sn[:, 1:2]))
    for name, val in zip(FIELDS, ref.structure(v, x, rbf1)):
        kat[f"snelson/{name}"] = val.numpy()
    kat["snelson/x"], kat["snelson/v"], kat["snelson/coeffs"] = x.numpy(), v.numpy(), rbf1.numpy()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)

    # ---- 3. seeded structure cases (small enough that the unmodified reference is exact) ---
    cases = {
        "d3": (500, 3, 2, rbf1, 11, "randn", 1.0),
        "d8": (400, 8, 4, rbf1, 12, "randn", 1.0),
        "d8_dense": (600, 8, 3, rbf1, 13, "rand", 1.0),
        "d11_r2": (250, 11, 2, mat15_2, 14, "randn", 1.0),
        "d18": (150, 18, 3, rbf1, 15, "randn", 1.0),
        "d24_r3": (100, 24, 2, mat15_3, 16, "randn", 1.0),
        "d5_r2": (700, 5, 5, rbf2, 17, "randn", 2.0),
        "d1_r0": (64, 1, 2, torch.tensor([1.0]), 18, "randn", 1.0),
    }
    st = {}
    for name, (N, d, L, c, seed, dist, scale) in cases.items():
        x, v = inputs(N, d, L, seed, dist, scale)
        res = ref.structure(v, x, c)
        assert res[4].shape[0] < 16000, "case too large for the unmodified reference table"
        assert torch.equal(res[7], ref.filter(v, x, c))
        st[f"{name}/x"], st[f"{name}/v"], st[f"{name}/coeffs"] = x.numpy(), v.numpy(), c.numpy()
        for fname, val in zip(FIELDS, res):
            st[f"{name}/{fname}"] = val.numpy()
    np.savez_compressed(os.path.join(HERE, "structure.npz"), **st)

    # ---- 4. autograd op ---------------------------------------------------------------------
    ag = {}
    with contextlib.redirect_stdout(io.StringIO()):
        dks = {
            "rbf1": bk.DiscretizedKernelFN(bk.rbf, 1),
            "rbf2": bk.DiscretizedKernelFN(bk.rbf, 2),
            "mat15_2": bk.DiscretizedKernelFN(lambda d2: bk.Matern.apply(d2, 1.5), 2),
            "mat25_1": bk.DiscretizedKernelFN(lambda d2: bk.Matern.apply(d2, 2.5), 1),
        }
    for name, (N, d, L, kern, seed) in {
        "rbf1_d3": (300, 3, 4, "rbf1", 21),
        "rbf2_d8": (200, 8, 2, "rbf2", 22),
        "mat15_d5": (250, 5, 3, "mat15_2", 23),
        "mat25_d2": (400, 2, 1, "mat25_1", 24),
        "rbf1_d1": (200, 1, 5, "rbf1", 25),
    }.items():
        x, v = inputs(N, d, L, seed)
        g = torch.Generator().manual_seed(seed + 1000)
        go = torch.randn(N, L, generator=g)
        xr = x.clone().requires_grad_(True)
        vr = v.clone().requires_grad_(True)
        out = bk.LatticeFilterGeneral.apply(vr, xr, dks[kern])
        out.backward(go)
        # case A of backward (only the source needs a gradient)
        vr2 = v.clone().requires_grad_(True)
        out2 = bk.LatticeFilterGeneral.apply(vr2, x, dks[kern])
        out2.backward(go)
        ag.update({f"{name}/x": x.numpy(), f"{name}/v": v.numpy(), f"{name}/grad_out": go.numpy(),
                   f"{name}/out": out.detach().numpy(), f"{name}/grad_src": vr.grad.numpy(),
                   f"{name}/grad_ref": xr.grad.numpy(), f"{name}/grad_src_only": vr2.grad.numpy(),
                   f"{name}/coeffs": dks[kern].get_coeffs().numpy(),
                   f"{name}/deriv_coeffs": dks[kern].get_deriv_coeffs().numpy()})
    np.savez_compressed(os.path.join(HERE, "autograd.npz"), **ag)

    # ---- 5. beyond the reference's first table doubling (M > 16383) ---------------------------
    large = {}
    for name, (N, d, L, c, seed) in {
        "d8_n20000": (20000, 8, 4, rbf1, 31),
        "d11_r2_n6000": (6000, 11, 2, mat15_2, 32),
        "d18_n3000": (3000, 18, 3, rbf1, 33),
    }.items():
        x, v = inputs(N, d, L, seed)
        fx = ref_fixed.structure(v, x, c)
        un = ref.structure(v, x, c)
        assert fx[4].shape[0] > 16383
        rec = {"N": N, "d": d, "L": L, "seed": seed, "coeffs": c.tolist(), "M": int(fx[4].shape[0]),
               "x_sha256": sha(x.numpy()), "v_sha256": sha(v.numpy()),
               "sha256": {fname: sha(val.numpy()) for fname, val in zip(FIELDS, fx)}}
        of, ou = fx[7].numpy().astype(np.float64), un[7].numpy().astype(np.float64)
        rec["unmodified_reference"] = {
            "M": int(un[4].shape[0]),
            "sha256": {fname: sha(val.numpy()) for fname, val in zip(FIELDS, un)},
            "out_rel_l2_vs_fixed": float(np.linalg.norm(of - ou) / np.linalg.norm(of)),
            "out_rows_differing": int((np.abs(of - ou).max(axis=1) > 0).sum()),
            "greedy_equal": bool(torch.equal(fx[0], un[0])), "rank_equal": bool(torch.equal(fx[1], un[1])),
            "weights_equal": bool(torch.equal(fx[3], un[3])),
            "offsets_differing": int((fx[2] != un[2]).sum()),
        }
        large[name] = rec
    with open(os.path.join(HERE, "large.json"), "w") as f:
        json.dump(large, f, indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Round-2 experiment driver: one-shot kernels vs TMA-ring kernels for splat / slice at a BASELINE.json shape.

    python profiles/exp_ring.py [--workload A] [--reps 30]

Prints one JSON line per variant (CUDA-event times on the launching stream, V/out rotating over 4 buffer pairs) and
the largest relative difference between the two forms' results.  Tuning knobs are environment variables read by the
library on every call (SGP_RING, SGP_SPLAT_STAGES, SGP_SLICE_STAGES, SGP_SLICE_PASSES, SGP_RING_OCC)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_gp_b200 as sg  # noqa: E402
from simplex_gp_b200 import _capi  # noqa: E402
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr  # noqa: E402
import bench  # noqa: E402


def timed(fn, reps, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="A")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    N, d, L = w["N"], w["d"], w["L"]
    coeffs = bench.COEFFS[(w["kernel"], w["order"])]
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, d, generator=g).to(dev)
    lat = sg.Lattice(x, coeffs)
    M = lat.M
    lib = _capi.lib()
    st = _stream_ptr(dev)
    gv = torch.Generator(device=dev).manual_seed(1)
    Vs = [torch.randn(N, L, generator=gv, device=dev) for _ in range(4)]
    outs = [torch.empty(N, L, device=dev) for _ in range(4)]
    Lv = (L + 3) // 4 * 4 if L > 4 else L
    buf0 = torch.empty(M, Lv, device=dev)
    buf1 = torch.empty(M, Lv, device=dev)
    rows = lat.rows
    v_out = lat._view(lat._table(False, True), None, lat.exact)

    def splat(i):
        V = Vs[i % 4]
        _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(V), V.stride(0), L,
                                       _ptr(buf0), Lv, st))

    def slice_(i):
        o = outs[i % 4]
        _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(buf1), Lv, _ptr(o), o.stride(0), L, st))

    def mvm(i):
        lat.mvm(Vs[i % 4], out=outs[i % 4])

    def env(**kw):
        for k, v in kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    res = {}
    # correctness: ring vs one-shot
    env(SGP_RING=0)
    splat(0); ref_splat = buf0.clone()
    buf1.copy_(torch.randn(M, Lv, generator=gv, device=dev))
    slice_(0); ref_slice = outs[0].clone()
    mvm(1); ref_mvm = outs[1].clone()
    env(SGP_RING=1)
    buf0.fill_(float("nan"))
    splat(0)
    res["splat_rel"] = float((buf0 - ref_splat).norm() / ref_splat.norm())
    res["splat_nan"] = int(torch.isnan(buf0).sum())
    slice_(0)
    res["slice_rel"] = float((outs[0] - ref_slice).norm() / ref_slice.norm())
    mvm(1)
    res["mvm_rel"] = float((outs[1] - ref_mvm).norm() / ref_mvm.norm())
    env(SGP_SPLAT_SCAN=1)
    buf0.fill_(float("nan"))
    splat(0)
    res["splat_scan_rel"] = float((buf0 - ref_splat).norm() / ref_splat.norm())
    res["splat_scan_nan"] = int(torch.isnan(buf0).sum())
    env(SGP_SPLAT_SCAN=None)
    print(json.dumps({"workload": args.workload, "M": M, "check": res}), flush=True)

    def report(name, **kw):
        print(json.dumps({"variant": name, **{k: round(v, 2) for k, v in kw.items()}}), flush=True)

    env(SGP_RING=0)
    report("one-shot", splat_us=timed(splat, args.reps), slice_us=timed(slice_, args.reps), mvm_us=timed(mvm, args.reps))
    env(SGP_RING=1)
    report("ring default", splat_us=timed(splat, args.reps), slice_us=timed(slice_, args.reps), mvm_us=timed(mvm, args.reps))
    if args.sweep:
        for scan in (0, 1):
            for stg in (2, 3):
                for occ in (0, 3, 4):
                    env(SGP_SPLAT_SCAN=scan, SGP_SPLAT_STAGES=stg, SGP_RING_OCC=occ)
                    try:
                        report(f"splat scan={scan} stages={stg} occ={occ}", splat_us=timed(splat, args.reps))
                    except Exception as exc:
                        print(json.dumps({"variant": f"splat scan={scan} stages={stg} occ={occ}", "error": str(exc)}), flush=True)
        env(SGP_SPLAT_SCAN=None, SGP_SPLAT_STAGES=None, SGP_RING_OCC=None)
        for stg in (2, 3):
            for ps in (1, 2, 4):
                env(SGP_SLICE_STAGES=stg, SGP_SLICE_PASSES=ps)
                try:
                    report(f"slice stages={stg} passes={ps}", slice_us=timed(slice_, args.reps))
                except Exception as exc:
                    print(json.dumps({"variant": f"slice stages={stg} passes={ps}", "error": str(exc)}), flush=True)
        env(SGP_SLICE_STAGES=None, SGP_SLICE_PASSES=None, SGP_RING_OCC=None)
    # graph replay of the whole product, both forms
    for rs in (0, 1):
        env(SGP_RING=1, SGP_RING_SPLAT=rs, SGP_RING_SLICE=1, SGP_SPLAT_PREFETCH=None, SGP_GRAPH_ZERO_AFTER=None)
        report(f"splat ring={rs}", splat_us=timed(splat, args.reps))
        graphs = [lat.capture(Vs[k], outs[k]) for k in range(4)]
        for k in range(4):
            graphs[k].replay()
        torch.cuda.synchronize()
        ok = float((outs[1] - ref_mvm).norm() / ref_mvm.norm())
        report(f"graph splat ring={rs}", mvm_us=timed(lambda i: graphs[i % 4].replay(), max(args.reps, 300), warm=20), rel=ok * 1e6)
        del graphs


if __name__ == "__main__":
    main()

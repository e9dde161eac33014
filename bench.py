#!/usr/bin/env python
"""Benchmark of the lattice MVM hot path (BASELINE.json metric): lattice MVM/s at N=1M, d=8, 16 RHS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one MVM (splat -> (d+1) x blur -> slice) of an [N, 16] RHS block on a PRE-BUILT lattice, inputs
resident in HBM (north star: the lattice is built once per hyper-parameter step and reused by every CG/Lanczos
MVM).  At N GPUs the lattice is built on rank 0, NCCL-broadcast, and every rank filters its own 16-column block
(weak scaling in RHS blocks, no data-path collective); `value` = blocks filtered per second over all ranks.

Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lattice MVM/s (N=1M,d=8,16 RHS)"   # the metric is quoted on workload A; other workloads label themselves
UNIT = "MVM/s"
WORKLOADS = {
    # name: (N, d, L, kernel, order)   -- SURVEY.md section 8: config A is the metric's configuration
    "A": dict(N=1_000_000, d=8, L=16, kernel="rbf", order=1),
    # the other BASELINE.json configurations: parity-test cases, timed for information only (--workload)
    "A1": dict(N=1_000_000, d=8, L=1, kernel="rbf", order=1),   # configs[1]: the single-RHS MVM
    "A11": dict(N=1_000_000, d=8, L=11, kernel="rbf", order=1),  # the training step's block [y | 10 probes] on the metric lattice
    "A12": dict(N=1_000_000, d=8, L=12, kernel="rbf", order=1),  # ... as the solver pads it (16-byte vectors)
    "B": dict(N=16_600, d=18, L=11, kernel="rbf", order=1),
    "C": dict(N=2_050_000, d=11, L=16, kernel="matern1.5", order=2),
    "D10": dict(N=1_000_000, d=24, L=4, kernel="matern1.5", order=3),
    "D": dict(N=10_000_000, d=24, L=1, kernel="matern1.5", order=3, build_nbr=False),   # ~122 GB on one GPU
}
COEFFS = {   # tests/golden/coeffs.json (the reference's DiscretizedKernelFN)
    ("rbf", 1): [0.34608543, 1.0, 0.34608543],
    ("matern1.5", 2): [0.15233751, 0.50067621, 1.0, 0.50067621, 0.15233751],
    ("matern1.5", 3): [0.08435782, 0.24239115, 0.60311586, 1.0, 0.60311586, 0.24239115, 0.08435782],
}
RBF1 = [0.34608543, 1.0, 0.34608543]   # get_coeffs(rbf, 1), tests/golden/coeffs.json


def kernel_source_stamp():
    """sha256 over the CUDA sources and the public header: profiles/traffic.json is only valid for the kernels it was
    captured from (profiles/summarize.py --traffic writes the same stamp)."""
    import hashlib
    h = hashlib.sha256()
    files = []
    for base in (os.path.join(ROOT, "simplex-gp_b200", "csrc"), os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(base)):
            if f.endswith((".cu", ".cuh", ".h")):
                files.append(os.path.join(base, f))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def load_traffic(workload="A"):
    """{kernel name: dram__bytes_read.sum + dram__bytes_write.sum per launch} from the committed `ncu --set full` capture
    (profiles/traffic.json, cold-cache replays of profiles/ncu_mvm.py), or {} when the capture is older than the
    kernels: a stale figure is reported as null, never as a number."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return {}, "no profiles/traffic.json"
    if t.get("source_stamp") != kernel_source_stamp():
        return {}, f"profiles/traffic.json is stale (captured for sources {t.get('source_stamp')})"
    if t.get("workload", "A") != workload:
        return {}, f"profiles/traffic.json was captured at workload {t.get('workload', 'A')}, not {workload}"
    return t.get("kernels", {}), f"profiles/traffic.json ({t.get('report')})"


def workload_name(w):
    return f"N={w['N']},d={w['d']},L={w['L']},{w['kernel']} order {w['order']},x~N(0,I),lengthscale 1"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU filter (oracle/_ref when built, else the C port)
# ---------------------------------------------------------------------------------------------
def cpu_reference_filter():
    """Returns (callable(src, ref, coeffs) -> out, kind)."""
    import torch
    from oracle import build_oracle
    mod = build_oracle.load_ref(fixed=False)
    if mod is not None:
        return (lambda s, r, c: mod.filter(s, r, c)), "reference"
    from oracle import oracle as port

    def run(s, r, c):
        return torch.from_numpy(port.filter(s.numpy(), r.numpy(), c.numpy()))
    return run, "port"


def time_cpu_filter(w, n_sample, repeats):
    """Best-of-`repeats` wall time of the reference filter on the first n_sample points of the workload."""
    import torch
    fn, kind = cpu_reference_filter()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_sample, w["d"], generator=g)
    v = torch.randn(n_sample, w["L"], generator=g)
    c = torch.tensor(COEFFS[(w["kernel"], w["order"])])
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn(v, x, c)
        best = min(best, time.perf_counter() - t0)
    return best, kind


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(1)
    # One reference call at the full workload costs 2-7 s of one core (the code is single-threaded), so the run is
    # bounded in STEPS, not in points: every executed step is a full-size filter() call (a 100k-point sample scaled
    # linearly was tried and is 57 % pessimistic on the GPU box's CPU: per-point cost depends on M/N and cache size),
    # and at most 40 steps / 1 warm-up are executed however many were asked for -- the reference is deterministic CPU
    # code, more repetitions add nothing but minutes.
    n_sample = int(w["N"])
    steps, warm = max(1, min(args.steps, 40)), max(0, min(args.warmup, 1))
    fn, kind = cpu_reference_filter()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_sample, w["d"], generator=g)
    v = torch.randn(n_sample, w["L"], generator=g)
    c = torch.tensor(COEFFS[(w["kernel"], w["order"])])
    for _ in range(warm):
        fn(v, x, c)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(v, x, c)
    dt = time.perf_counter() - t0
    per_step = dt / steps
    # scale the sample linearly to the full N: one full-size MVM costs (N / n_sample) sample filters
    value = 1.0 / (per_step * (w["N"] / n_sample))
    sample = (f"{steps} reference filter() calls executed ({args.steps} requested), lattice rebuilt inside each call as the "
              f"reference does, on all {n_sample} points (the full workload in every step); single-threaded code, "
              f"{os.cpu_count()} host cores present")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": per_step * 1e3 * (w["N"] / n_sample), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(w), "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# host placement: one rank per GPU, its threads and pinned buffers on the GPU's NUMA node
# ---------------------------------------------------------------------------------------------
def pin_rank_to_gpu_numa(local: int):
    """Restrict this process to the CPUs local to GPU `local` (sysfs local_cpulist of its PCI device), so that the
    pinned staging buffers it allocates afterwards are first-touched on that NUMA node: in round 1 all 8 ranks ran on
    CPUs 0-31 and the end-to-end filter() scaled 3.0x at 8 GPUs.  Returns a short description for the bench line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return f"gpu {local}: no local cpus within the allowed set"
        os.sched_setaffinity(0, cpus)
        node = None
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        return f"gpu {local} ({bus}) numa {node}: {len(cpus)} cpus {spec}"
    except Exception as exc:
        return f"not pinned: {exc!r}"


def timed_steps(step, steps, warm, world, dev):
    """`warm` untimed + exactly `steps` timed calls of step(i): barrier + synchronize on both sides, CUDA events on the
    launching stream, MAX over ranks.  Returns elapsed milliseconds."""
    import torch
    import torch.distributed as dist
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def array_checksum(t):
    """Order-sensitive 64-bit checksum of a tensor's bytes (for comparing broadcast / merged lattices across ranks)."""
    import torch
    b = t.contiguous().view(-1).view(torch.uint8)
    pad = (-b.numel()) % 8
    if pad:
        b = torch.cat([b, torch.zeros(pad, dtype=torch.uint8, device=b.device)])
    w = b.view(torch.int64)
    idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64)
    return int(((w ^ (idx * 0x1E3779B97F4A7C15)).sum()).item())


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def strong_and_check(args, w, lat, world, rank, dev, ms_full_block, mode, blur_arg, steps, warm):
    """N > 1, inside the line the driver records:
    strong -- ONE L-column job (the block every rank filtered whole in the weak run) with its columns split over the
              ranks: MVM/s of that job and the speed-up over one GPU filtering all L columns (this run's own weak step);
    check  -- NCCL correctness against the single-GPU result: the broadcast lattice is rank 0's (checksums of replay /
              keys / nbr gathered from every rank), the column-sharded product gathered with all_gather equals the
              unsharded one, and the point-sharded product (sharded build + all-reduce of the lattice values) equals it
              too, with the sharded build's keys equal to rank 0's."""
    import torch
    import torch.distributed as dist

    from simplex_gp_b200.distributed import ColumnShardedOperator, PointShardedLattice, shard_columns, shard_points
    N, d, L = w["N"], w["d"], w["L"]
    coeffs = COEFFS[(w["kernel"], w["order"])]
    lo, hi = shard_columns(L, world, rank)
    Lr = hi - lo
    strong = None
    if Lr >= 1 and L >= world:
        gv = torch.Generator(device=dev).manual_seed(4321)      # the same job on every rank
        n_rot = 4
        Vfull = [torch.randn(N, L, generator=gv, device=dev) for _ in range(n_rot)]
        Vs = [v[:, lo:hi].contiguous() for v in Vfull]
        del Vfull
        outs = [torch.empty(N, Lr, device=dev) for _ in range(n_rot)]
        graphs = None if args.no_graph else [lat.capture(Vs[k], outs[k], mode=mode, blur=blur_arg) for k in range(n_rot)]

        def step(i):
            if graphs is not None:
                graphs[i % n_rot].replay()
            else:
                lat.mvm(Vs[i % n_rot], out=outs[i % n_rot], mode=mode, blur=blur_arg)

        ms = timed_steps(step, steps, warm, world, dev) / steps
        strong = {"job": f"one {L}-column block, columns split {world} ways ({Lr} per rank)", "value": 1e3 / ms, "unit": UNIT,
                  "ms_per_step": ms, "one_gpu_ms_per_step": ms_full_block, "speedup_vs_1": ms_full_block / ms,
                  "limiter": "the per-point index streams (replay / row-sorted entries: 8.5-16 B per point-vertex) do not "
                             "shrink with the column count; column sharding pays when M*L >> N*(d+1) (workload C)"}
        del graphs, Vs, outs
    # ---- correctness ----------------------------------------------------------------------------------------------
    sums = torch.tensor([array_checksum(lat.replay), array_checksum(lat.keys), array_checksum(lat.nbr)], device=dev)
    allsums = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    same = all(bool(torch.equal(a, allsums[0])) for a in allsums)
    gv = torch.Generator(device=dev).manual_seed(777)
    V = torch.randn(N, L, generator=gv, device=dev)             # identical on every rank
    want = lat.mvm(V).clone()
    op = ColumnShardedOperator(lat)
    full = op.matmul_full(V)
    e_col = float((full - want).norm() / want.norm())
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(0)).to(dev)
    ps = PointShardedLattice(x, coeffs)      # first use: communicator channels for all_gather, allocator
    del ps
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    ps = PointShardedLattice(x, coeffs)
    torch.cuda.synchronize()
    ps_build_ms = (time.perf_counter() - t0) * 1e3
    plo, phi = shard_points(N, world, rank)
    mine = ps.mvm(V[plo:phi].contiguous())
    e_pt = torch.tensor([float((mine - want[plo:phi]).norm() / want[plo:phi].norm())], device=dev, dtype=torch.float64)
    dist.all_reduce(e_pt, op=dist.ReduceOp.MAX)
    keys_same = torch.tensor([int(ps.M == lat.M and array_checksum(ps.local.keys) == int(allsums[0][1].item()))], device=dev)
    dist.all_reduce(keys_same, op=dist.ReduceOp.MIN)
    e_colt = torch.tensor([e_col], device=dev, dtype=torch.float64)
    dist.all_reduce(e_colt, op=dist.ReduceOp.MAX)
    check = {"broadcast_lattice_identical_on_all_ranks": bool(same),
             "column_sharded_all_gather_rel_err": float(e_colt.item()),
             "point_sharded_rel_err": float(e_pt.item()),
             "point_sharded_build_keys_equal_rank0_build": bool(keys_same.item()),
             "point_sharded_build_ms": ps_build_ms, "tolerance": 1e-5,
             "ok": bool(same and float(e_colt.item()) < 1e-5 and float(e_pt.item()) < 1e-5 and bool(keys_same.item()))}
    del ps, V, want, full, mine, x
    return strong, check


def run_point_sharded(args, w, world, rank, dev, placement):
    """`--scaling point`: the POINTS are split over the ranks (BASELINE.json configs[4]).  Every rank builds the lattice
    of its own points only, the key lists are merged (sharded build), and a step is: local splat of the rank's rows into
    the full [M, L] lattice, NCCL all-reduce of M*L*4 bytes, blur (replicated), local slice.  `value` = whole-job MVM/s."""
    import torch
    import torch.distributed as dist

    from simplex_gp_b200.distributed import PointShardedLattice, shard_points
    N, d, L = w["N"], w["d"], w["L"]
    coeffs = COEFFS[(w["kernel"], w["order"])]
    steps, warm = args.steps, max(args.warmup, 3)
    lo, hi = shard_points(N, world, rank)
    g = torch.Generator().manual_seed(0)
    # every rank draws the same stream and keeps its share (the whole set is only 4*N*d bytes on the host)
    x_loc = torch.randn(N, d, generator=g)[lo:hi].contiguous().to(dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ps = PointShardedLattice(x_loc, coeffs, x_is_local=True, build_nbr=w.get("build_nbr", True))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    build_ms = (time.perf_counter() - t0) * 1e3
    M = ps.M
    n_rot = 2
    gv = torch.Generator(device=dev).manual_seed(1234 + rank)
    Vs = [torch.randn(hi - lo, L, generator=gv, device=dev) for _ in range(n_rot)]
    outs = [torch.empty(hi - lo, L, device=dev) for _ in range(n_rot)]

    col_blur = world > 1 and L % world == 0 and ps.local.groups is not None and args.column_blur

    def step(i):
        outs[i % n_rot] = ps.mvm(Vs[i % n_rot], out=outs[i % n_rot], column_blur=col_blur)

    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    elapsed_ms = timed_steps(step, steps, warm, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = elapsed_ms / steps
    # the exchange step alone
    vals = torch.zeros(M, (L + 3) // 4 * 4 if L > 4 else L, device=dev)

    def ar(i):
        if world > 1:
            dist.all_reduce(vals)

    ar_ms = timed_steps(ar, max(5, min(steps, 20)), 2, world, dev) / max(5, min(steps, 20)) if world > 1 else 0.0
    peak, peak_src = load_peaks()
    mem = torch.tensor([torch.cuda.max_memory_allocated(dev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    if rank == 0:
        r = w["order"]
        alg = 4 * (2 * N * L + 4 * N * (d + 1) + 2 * M * L + (d + 1) * (2 * M * L + 2 * r * M))
        line = {
            "metric": f"lattice MVM/s ({workload_name(w)})", "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(w), "M": M, "points_per_rank": hi - lo, "M_local_rank0": ps.M_local,
                       "sharding": "points split over the ranks; sharded lattice build (per-rank build + key-list merge); per step: "
                                   + ("local splat -> NCCL reduce-scatter over column blocks -> blur of L/ranks columns -> "
                                      "all-gather -> local slice (the same M*L*4 bytes on the wire as an all-reduce)" if col_blur
                                      else "local splat -> NCCL all-reduce of M*L*4 bytes -> replicated blur -> local slice"),
                       "blur": "column-sharded" if col_blur else "replicated",
                       "l2": "inputs larger than L2 (index tables and lattice values exceed 126 MB)"},
            "clocks": clocks, "gpu_launches": steps * (2 + (len(ps.local.groups["list"]) if ps.local.groups else d + 1)),
            "lattice_build_ms": build_ms, "allreduce_bytes_per_step": int(vals.numel() * 4), "allreduce_ms": ar_ms,
            "allreduce_share_of_step": ar_ms / ms_per_step if ms_per_step else None,
            "peak_device_memory_gb": float(mem.item()) / 1e9,
            "mvm_roofline": {"alg_bytes": alg, "achieved": alg / ms_per_step / 1e6, "peak": peak * world, "unit": "GB/s",
                             "frac": alg / ms_per_step / 1e6 / (peak * world), "peak_source": peak_src,
                             "note": "whole-job algorithmic bytes over the aggregate HBM peak of the ranks; the blur's "
                                     "(d+1)(2ML+2rM) bytes are replicated on every rank, so they do not speed up"},
            "limiter": ("the exchange (reduce-scatter + all-gather of the lattice values and their two re-layout copies) and the "
                        "narrow-row blur: its time falls less than linearly with the column count") if col_blur else
                       ("the replicated blur (every rank moves the whole lattice d+1 times) plus the all-reduce; only "
                        "splat and slice shrink with the rank count"),
            "host_placement": placement,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import simplex_gp_b200 as sg
    from simplex_gp_b200 import _capi
    from simplex_gp_b200.distributed import broadcast_lattice
    from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = pin_rank_to_gpu_numa(local) if not args.no_pin else "not pinned (--no-pin)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    if args.scaling == "point":
        return run_point_sharded(args, w, world, rank, dev, placement)

    N, d, L = w["N"], w["d"], w["L"]
    L_job = L
    if args.scaling == "strong":   # the job is ONE L-column block; every rank filters L/world of its columns
        from simplex_gp_b200.distributed import shard_columns
        lo_c, hi_c = shard_columns(L, world, rank)
        L = max(1, hi_c - lo_c)
    coeffs = COEFFS[(w["kernel"], w["order"])]
    steps, warm = args.steps, max(args.warmup, 3)
    n_rot = 4   # V/out buffer pairs rotated so consecutive steps never re-read the same RHS from L2

    # --- lattice: built once on rank 0 and broadcast (north star) --------------------------------
    g = torch.Generator().manual_seed(0)
    x_host = torch.randn(N, d, generator=g)
    build_ms = bcast_ms = None
    lat = None
    if rank == 0:
        x = x_host.to(dev)
        lkw = {"build_nbr": w.get("build_nbr", True)}
        if N <= 2_500_000:
            # warm-up: allocator and module load, then the first build that sizes its tables from the previous lattice of
            # the shape (what every hyper-parameter step after the first does): the timed build is a steady-state one
            for _ in range(2):
                sg.Lattice(x, coeffs, **lkw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = sg.Lattice(x, coeffs, **lkw)
        e1.record()
        torch.cuda.synchronize()
        build_ms = e0.elapsed_time(e1)
    bcast_cold_ms = None
    if world > 1:
        # cold: first collective of the process (communicator set-up); warm: what a hyper-parameter step pays -- header +
        # one byte buffer over NVLink, then the derived tables (blur groups, row-sorted entries) rebuilt on each rank
        lat0 = lat
        for attempt in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lat = broadcast_lattice(lat0, src=0, device=dev)
            torch.cuda.synchronize()
            dist.barrier()
            dt = (time.perf_counter() - t0) * 1e3
            if attempt == 0:
                bcast_cold_ms = dt
                if rank != 0:
                    del lat
            else:
                bcast_ms = dt
        del lat0
    M = lat.M

    # --- this rank's RHS blocks -----------------------------------------------------------------------
    gv = torch.Generator(device=dev).manual_seed(1234 + rank)
    Vs = [torch.randn(N, L, generator=gv, device=dev) for _ in range(n_rot)]
    outs = [torch.empty(N, L, device=dev) for _ in range(n_rot)]
    mode = {"atomic": _capi.MODE_ATOMIC, "gather": _capi.MODE_GATHER, "tiles": _capi.MODE_TILES,
            "rows": _capi.MODE_ROWS, "auto": _capi.MODE_AUTO}[args.splat]
    if mode == _capi.MODE_AUTO:
        mode = _capi.MODE_ROWS if lat.rows is not None else _capi.MODE_ATOMIC
    if mode == _capi.MODE_GATHER and lat.csr_ptr is None:
        lat._build_csr()
    if mode == _capi.MODE_TILES and lat.tiles is None:
        lat._build_tiles()
    use_groups = lat.groups is not None and args.blur != "axis"

    blur_arg = "groups" if use_groups else "axis"
    graphs = None
    if not args.no_graph:   # one CUDA graph per buffer pair: a step is one replay (memset + 5 kernel nodes)
        graphs = [lat.capture(Vs[k], outs[k], mode=mode, blur=blur_arg) for k in range(n_rot)]

    def step(i):
        if graphs is not None:
            graphs[i % n_rot].replay()
        else:
            lat.mvm(Vs[i % n_rot], out=outs[i % n_rot], mode=mode, blur=blur_arg)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # sampled over warm-up + timed region (a 50-step timed region alone lasts only ~11 ms)
    elapsed_ms = timed_steps(step, steps, warm, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = elapsed_ms / steps
    value = (1 if args.scaling == "strong" else world) * steps / (elapsed_ms * 1e-3)

    # --- N > 1: the same 16-column job split over the ranks (strong), and NCCL correctness in the line -------------
    strong = check = None
    if world > 1 and args.scaling == "weak" and not args.no_multi_gpu_check:
        strong, check = strong_and_check(args, w, lat, world, rank, dev, ms_per_step, mode, blur_arg, steps, warm)

    # --- per-stage device times (CUDA events on the launching stream), same buffers, rank 0 -----------
    peak, peak_src = load_peaks()
    traffic, traffic_src = load_traffic(args.workload if L == w["L"] else None)
    roofline = stages = None
    if rank == 0:
        lib = _capi.lib()
        # as Lattice.mvm: lattice rows padded to a multiple of four channels, ragged right-hand sides through a
        # zero-padded copy where that pays
        Lv = lat.lattice_width(L) if mode == _capi.MODE_ROWS else L
        buf0, buf1 = lat._scratch(Lv)
        pad_src = mode == _capi.MODE_ROWS and L > 4 and L % 4 != 0 and lat._pads_ragged_src()
        Vs_pad = None
        if pad_src:
            Vs_pad = [torch.zeros((N, (L + 3) // 4 * 4), dtype=torch.float32, device=dev) for _ in Vs]
            for vp_, v_ in zip(Vs_pad, Vs):
                vp_[:, :L].copy_(v_)
        cnp = lat.coeffs
        st = _stream_ptr(dev)
        reps = max(5, min(steps, 20))
        where = C.c_int(0)
        fast = 0 if lat.exact else 1
        v_in = lat._view(lat._table(False, False), None, lat.exact)
        v_axis = lat._view(exact=lat.exact)
        v_out = lat._slice_view(Lv, use_groups, lat.exact)
        tv_in = lat._tiles_view(False) if mode == _capi.MODE_TILES else None
        tv_out = lat._tiles_view(use_groups) if mode == _capi.MODE_TILES else None
        garr = lat.groups["array"] if use_groups else None
        # Each stage is launched `reps` times back to back between two events (after two warm-up launches), so that the
        # average is the kernel's device time and not the launch latency of an idle stream; V/out rotate as in the step.
        def splat_stage(i):
            V = Vs_pad[i % n_rot] if pad_src else Vs[i % n_rot]
            if mode == _capi.MODE_TILES:
                _capi.check(lib.sgp_splat_tiles(C.byref(tv_in), _ptr(V), V.stride(0), L, _ptr(buf0), st))
            elif mode == _capi.MODE_ROWS:
                # the kernel alone (its memset is timed apart: in the MVM's graph it runs beside the previous slice);
                # repeated launches keep accumulating into buf0, which is irrelevant for the timing
                _capi.check(lib.sgp_mvm_stage_splat_prezeroed(_ptr(lat.rows["ent"]), _ptr(lat.rows["seg_row"]), lat.rows["n"], N,
                                                              M, _ptr(V), V.stride(0), int(V.shape[1]), _ptr(buf0), Lv, st))
            else:
                _capi.check(lib.sgp_splat(C.byref(v_in if mode == _capi.MODE_ATOMIC else v_axis), _ptr(V), V.stride(0), L,
                                          _ptr(buf0), mode, st))

        def blur_stage(i):
            if use_groups:
                _capi.check(lib.sgp_blur_groups(garr, len(garr), M, lat.order, _fp(cnp), cnp.shape[0], Lv, _ptr(buf0),
                                                _ptr(buf1), C.byref(where), fast, st))
            else:
                _capi.check(lib.sgp_blur(C.byref(v_axis), _fp(cnp), cnp.shape[0], Lv, _ptr(buf0), _ptr(buf1),
                                         C.byref(where), st))

        def slice_stage(i):
            out = outs[i % n_rot]
            res_buf = buf1 if where.value else buf0
            if mode == _capi.MODE_TILES:
                _capi.check(lib.sgp_slice_tiles(C.byref(tv_out), _ptr(res_buf), L, _ptr(out), out.stride(0), fast, st))
            else:
                _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(res_buf), Lv, _ptr(out), out.stride(0), L, st))

        def stage_ms(fn):
            for i in range(2):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                fn(i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        buf0.zero_()
        t_splat = stage_ms(splat_stage)
        memset_ms = stage_ms(lambda i: buf0.zero_()) if mode == _capi.MODE_ROWS else None
        buf0.normal_()
        t_blur = stage_ms(blur_stage)
        buf0.normal_(); buf1.normal_()
        t_slice = stage_ms(slice_stage)
        r = lat.order
        n_blur = len(lat.groups["list"]) if use_groups else d + 1
        b_splat = 4 * (N * L + 2 * N * (d + 1) + M * L)
        b_blur = (d + 1) * 4 * (2 * M * L + 2 * r * M)
        b_slice = 4 * (M * L + 2 * N * (d + 1) + N * L)
        splat_kernel = {_capi.MODE_ATOMIC: "sgp_splat_atomic_kernel", _capi.MODE_GATHER: "sgp_splat_gather_kernel",
                        _capi.MODE_TILES: "sgp_splat_tiles_kernel",
                        # the selection rule of csrc/sgp_tiles.cu::splat_rows_impl
                        _capi.MODE_ROWS: "sgp_splat_ring_kernel" if (lib.sgp_ring_splat_enabled() and lat.rows["n"] >= 8 * M
                                                                       and lat.rows["n"] >= (1 << 21) and Lv % 4 == 0
                                                                       and 8 <= Lv <= 64) else "sgp_splat_rows_kernel"}[mode]
        stages = {
            "splat": {"ms": t_splat, "launches": 1, "alg_bytes": b_splat, "gbs": b_splat / t_splat / 1e6,
                      "kernel": splat_kernel, "memset_ms_timed_apart": memset_ms},
            "blur": {"ms": t_blur, "launches": n_blur, "alg_bytes": b_blur, "gbs": b_blur / t_blur / 1e6,
                     "kernel": "sgp_blur_group_kernel" if use_groups else "sgp_blur_kernel"},
            "slice": {"ms": t_slice, "launches": 1, "alg_bytes": b_slice, "gbs": b_slice / t_slice / 1e6,
                      "kernel": "sgp_slice_tiles_kernel" if mode == _capi.MODE_TILES else
                                ("sgp_slice_ring_kernel" if (lib.sgp_ring_slice_enabled() and
                                                            lib.sgp_slice_ring_supported(C.byref(v_out), _ptr(buf0), Lv))
                                 else "sgp_slice_kernel")},
        }
        dom = max(stages, key=lambda k: stages[k]["ms"] / stages[k]["launches"])
        per_launch_bytes = stages[dom]["alg_bytes"] / stages[dom]["launches"]
        per_launch_ms = stages[dom]["ms"] / stages[dom]["launches"]
        achieved = per_launch_bytes / per_launch_ms / 1e6
        roofline = {"bound": "hbm", "kernel": stages[dom]["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic.get(stages[dom]["kernel"]), "traffic_source": traffic_src,
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                    "share_of_step": stages[dom]["ms"] / ms_per_step}
        n_launches = 1 + n_blur + 1

    # --- end-to-end through the reference-facing call, host buffers ----------------------------------
    e2e = None
    if args.e2e_steps > 0:
        x_pin = x_host.pin_memory()
        gh = torch.Generator().manual_seed(99 + rank)
        v_pin = torch.randn(N, L, generator=gh).pin_memory()
        c_t = torch.tensor(coeffs)
        for _ in range(3):   # warm-up: module load, pinned-buffer pool (two output blocks alternate), allocator
            res = sg.filter(v_pin, x_pin, c_t, device=dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = sg.filter(v_pin, x_pin, c_t, device=dev)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": 4 * (N * L + N * d), "d2h_bytes_per_step": 4 * N * L,
               "call": "simplex_gp_b200.filter(src, ref, coeffs) with pinned host tensors = ONE sgp_filter_host call: H2D, "
                       "lattice build, MVM (atomic splat, per-axis blur, ring slice: one product per lattice), D2H",
               "steps": args.e2e_steps, "ms_per_call": float(dt.item()) / args.e2e_steps * 1e3,
               "checksum": float(res.double().sum().item())}
        if world > 1:
            # the same loop with the odd ranks started half a call late: ranks that run in lockstep all upload and then
            # all download, using one PCIe direction at a time; independent processes in production are not in lockstep
            dist.barrier()
            if rank % 2:
                time.sleep(e2e["ms_per_call"] * 0.5e-3)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                res = sg.filter(v_pin, x_pin, c_t, device=dev)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e["value_staggered_start"] = world * args.e2e_steps / float(dt.item())
        # where a call's time goes, each phase alone, MAX over ranks (all ranks run it at the same time, as in the call)
        x_dev, v_dev = x_pin.to(dev), v_pin.to(dev)
        out_pin = torch.empty(N, L).pin_memory()

        def phase(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            t = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def h2d():
            x_dev.copy_(x_pin, non_blocking=True)
            v_dev.copy_(v_pin, non_blocking=True)

        e2e["phases_ms"] = {
            "h2d": phase(h2d),
            "filter_on_device_tensors": phase(lambda: sg.filter(v_dev, x_dev, c_t)),
            "d2h": phase(lambda: out_pin.copy_(outs[0], non_blocking=True)),
        }
        e2e["host_placement"] = placement
        del x_dev, v_dev, out_pin

    # --- CPU baseline beside it (rank 0, N=1 only, bounded sample) ------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            t, kind = time_cpu_filter(w, N, repeats=2)
            cpu_baseline = {"value": 1.0 / t, "unit": UNIT, "cores": 1, "kind": kind,
                            "sample": f"best of 2 full reference filter() calls at N={N}, d={d}, L={L} (lattice build "
                                      f"included: the reference rebuilds it in every call); {os.cpu_count()} host cores "
                                      f"present, the reference code is single-threaded"}
        except Exception as exc:  # the baseline is reported, never required
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        alg_bytes = lat.algorithmic_bytes(L)
        line = {
            "metric": METRIC if args.workload == "A" else f"lattice MVM/s ({workload_name(w)})", "value": value,
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(w), "M": M, "columns_per_rank": L, "columns_per_job": L_job * (1 if args.scaling == "strong" else world),
                       "path": {"splat": {1: "atomic scatter", 2: "ordered gather", 3: "tiles", 4: "row-sorted segmented gather"}[mode],
                                "blur": "groups through shared memory" if use_groups else "one launch per axis",
                                "arithmetic": "reference order (exact)" if lat.exact else "fused multiply-add",
                                "launch": "eager" if graphs is None else "CUDA graph replay (one graph per MVM)"},
                       "sharding": "lattice built on rank 0 + NCCL broadcast; one 16-column RHS block per rank",
                       "l2": (f"inputs larger than L2: a step streams {4 * (2 * N * L + 5 * N * (d + 1)) / 1e6:.0f} MB of RHS, "
                              f"output and index tables through the 126 MB L2, and V/out rotate over {n_rot} buffer pairs "
                              f"({n_rot * 8 * N * L / 1e6:.0f} MB), so no step re-reads its inputs from cache; only the "
                              f"{4 * M * L / 1e6:.0f} MB lattice-value buffers stay L2-resident, by design")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": steps * n_launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "mvm_roofline": {"alg_bytes": alg_bytes, "achieved": alg_bytes / ms_per_step / 1e6, "peak": peak,
                             "unit": "GB/s", "frac": alg_bytes / ms_per_step / 1e6 / peak},
            "stages": stages, "lattice_build_ms": build_ms, "lattice_broadcast_ms": bcast_ms,
            "lattice_broadcast_cold_ms": bcast_cold_ms, "strong": strong, "multi_gpu_check": check,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A", choices=sorted(WORKLOADS))
    ap.add_argument("--splat", default="auto", choices=["auto", "rows", "tiles", "atomic", "gather"],
                    help="splat form: row-sorted segmented gather (default), locality tiles, atomic scatter, ordered gather")
    ap.add_argument("--blur", default="groups", choices=["groups", "axis"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong", "point"],
                    help="weak: one L-column RHS block per rank (default, plus a `strong` block in the line when N > 1); "
                         "strong: ONE L-column block split over the ranks; point: the POINTS split over the ranks "
                         "(sharded lattice build, all-reduce of the lattice values between splat and blur)")
    ap.add_argument("--column-blur", action="store_true",
                    help="--scaling point: reduce-scatter / column-sharded blur / all-gather instead of all-reduce + replicated "
                         "blur (measured slower at D10 on 2 GPUs: 8.2 vs 6.2 ms)")
    ap.add_argument("--no-pin", action="store_true", help="do not restrict the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-multi-gpu-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the MVM kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)
    return run_ours(args, w)


if __name__ == "__main__":
    sys.exit(main())

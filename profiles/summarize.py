#!/usr/bin/env python
"""Turn an .ncu-rep (from gpurun_out/, scratch) into the small text summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/mvm_r1e.ncu-rep profiles/r1_mvm_full.txt
    python profiles/summarize.py --launches gpurun_out/r1_launches.csv profiles/r1_launches.txt
    python profiles/summarize.py --traffic gpurun_out/mvm_r2.ncu-rep          # -> profiles/traffic.json (read by bench.py)

Needs the `ncu` CLI (present in the build container; no GPU required to read a report).
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors.sum",
    "smsp__inst_executed_op_global_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "derived__lts__lts2xbar_bytes.sum.per_second", "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]} (cold-cache, serialised replays)\n")
        for r in data:
            f.write(f"\n== {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            for m in METRICS:
                if m in idx:
                    f.write(f"   {m:86s} {r[idx[m]]:>16s} {units[idx[m]]}\n")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            name = r[hdr.index("Kernel Name")].split("(")[0]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[-1])
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {path.split('/')[-1]}: per-kernel totals\n")
        f.write(f"{'kernel':64s} {'launches':>8s} {'total us':>12s} {'avg us':>10s}\n")
        for k, v in agg.items():
            f.write(f"{k[:64]:64s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.2f}\n")


def traffic(rep):
    """profiles/traffic.json: DRAM bytes per launch of every kernel in the report (the LAST launch of each kernel
    name), stamped with the hash of the kernel sources it was captured from; bench.py reports a figure only while the
    stamp matches the sources it runs."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kernels = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]]) * scale[units[idx[m]]]
        kernels[name] = tot
    out = {"report": os.path.basename(rep), "source_stamp": bench.kernel_source_stamp(), "workload": "A",
           "how": "ncu --set full --clock-control none python profiles/ncu_mvm.py; dram__bytes_read.sum + dram__bytes_write.sum per launch (cold cache)",
           "kernels": kernels}
    with open(os.path.join(root, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2])
    elif sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2])

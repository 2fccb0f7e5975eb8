#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  tools/ncu_summary.py launches <launches.csv>        -> per-kernel share of the profiled region
  tools/ncu_summary.py raw <report.ncu-rep> [regex]   -> key metrics of the first matching launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg ",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "smsp__average_warps_issue_stalled",
    "lts__t_bytes.sum ", "l1tex__t_bytes.sum ",
]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\s+", " ", d["Kernel Name"])[:100]
        v = float(d["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, d["Metric Unit"]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    print(f"# {path}: per-kernel device time in the profiled region (ncu, cold-cache, serialised)")
    print(f"{'time':>14s} unit launches  share  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:14.1f} {v[2]:4s} {v[0]:8d} {100 * v[1] / tot:6.2f}%  {k}")


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for vals in rows[2:]:
        if pattern and not re.search(pattern, vals[ki]):
            continue
        print(f"# {path}: {vals[ki][:120]}")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEYS):
                if k_ok(h):
                    print(f"{h:95s} {u:16s} {v}")
        break


def k_ok(h):
    return not any(s in h for s in (".max.", ".min.", ".sum.pct", ".sum.per", "realtime", "peak_sustained_elapsed.")) \
        or h.startswith("dram__bytes_read.sum.per_second")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/<tag>_launches.csv  > profiles/<tag>_launches.txt
  python scripts/ncu_summary.py kernel   gpurun_out/<tag>_sampler.ncu-rep > profiles/<tag>_sampler.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[vi].replace(",", "")) / 1e3
    all_us = sum(v[1] for v in tot.values())
    print(f"# {path}: {len(rows) - 1} launches, {all_us / 1e3:.3f} ms device time under ncu (cold-cache, serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {100 * v[1] / all_us:6.1f}%")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"# {path}\n# kernel: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                print(f"{h:88s} {u:16s} {v}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])

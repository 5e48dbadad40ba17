#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back into the small text files committed under profiles/.

    python tools/ncu_summary.py TAG [--launches gpurun_out/launches_TAG.csv] [--rep gpurun_out/prof_TAG.ncu-rep] [--rows N]

writes profiles/TAG_launch_share.txt   per-kernel launch count / total device time / share of the listed launches
       profiles/TAG_launches.csv       the launch list itself (id, kernel, grid, block, ns)
       profiles/TAG_ncu_full.csv       every raw metric of the `--set full` capture (metric, unit, value per launch)
       profiles/TAG_ncu_key.txt        the handful of counters DESIGN.md argues from + instructions per check row by
                                       source section (needs -lineinfo and --import-source on)
Reads the .ncu-rep with `ncu -i` (works without a GPU).
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEY = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
       "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    agg = collections.OrderedDict()
    with open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w") as f:
        f.write("id,kernel,grid,block,ns\n")
        for r in rows:
            name = re.sub(r"\(.*", "", r[4])
            f.write('%s,"%s","%s","%s",%s\n' % (r[0], name, r[8], r[7], r[14]))
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[14])
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launch_share.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (%d launches; cold-cache, serialised: compare shares)\n" % len(rows))
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-70s launches %4d  total %12.0f ns  share %5.1f%%\n" % (name[:70], n, ns, 100 * ns / tot))


def ncu_page(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout


def full(rep, tag, rows_per_launch, sections):
    raw = list(csv.reader(ncu_page(rep, "--page", "raw").splitlines()))
    hdr, units, vals = raw[0], raw[1], raw[2:]
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_full.csv"), "w") as f:
        f.write("metric,unit," + ",".join("launch%d" % (i + 1) for i in range(len(vals))) + "\n")
        for i, h in enumerate(hdr):
            f.write('%s,%s,%s\n' % (h, units[i], ",".join('"%s"' % v[i] for v in vals)))
    col = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_key.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on, kernel %s\n" % vals[0][col["Kernel Name"]])
        for k in KEY:
            if k in col:
                f.write("%-90s %-12s %s\n" % (k, units[col[k]], " ".join(v[col[k]] for v in vals)))
        if "dram__bytes_read.sum" in col:
            f.write("traffic per launch = dram__bytes_read.sum + dram__bytes_write.sum (units above)\n")
        src = list(csv.reader(ncu_page(rep, "--page", "source", "--print-source", "cuda,sass").splitlines()))
        cur, per_file, per_sec = None, collections.Counter(), collections.Counter()
        for r in src:
            if len(r) >= 2 and r[0] == "File Path":
                cur = os.path.basename(r[1])
                continue
            if len(r) < 8 or not r[0].isdigit():
                continue
            try:
                n = int(r[7])
            except ValueError:
                continue
            per_file[cur] += n
            for name, (fn, a, b) in sections.items():
                if cur == fn and a <= int(r[0]) <= b:
                    per_sec[name] += n
        if per_file and rows_per_launch:
            f.write("\nwarp instructions per check row (launch of %d rows), by source file of the line they are attributed to\n" % rows_per_launch)
            for k, n in per_file.most_common():
                f.write("  %-28s %9.1f\n" % (k, n / rows_per_launch))
            if per_sec:
                f.write("by section of the kernel source\n")
                for k, n in per_sec.items():
                    f.write("  %-28s %9.1f\n" % (k, n / rows_per_launch))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--rows", type=int, default=0, help="check rows per profiled launch (frames x M)")
    ap.add_argument("--section", action="append", default=[], help="name:file:first:last (source line range)")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    secs = collections.OrderedDict()
    for s in a.section:
        n, fn, lo, hi = s.split(":")
        secs[n] = (fn, int(lo), int(hi))
    if a.launches:
        launches(a.launches, a.tag)
    if a.rep:
        full(a.rep, a.tag, a.rows, secs)

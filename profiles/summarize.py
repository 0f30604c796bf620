#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

  python profiles/summarize.py launches <launches.csv> [name-filter]   per-kernel totals and shares
  python profiles/summarize.py rep <file.ncu-rep>                       key counters of every captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path, flt=None):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0]
            if flt and flt not in name:
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("# %s  (filter=%s)  total %.1f us" % (path, flt, tot))
    print("%-44s %6s %12s %8s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s %6d %12.1f %7.1f%% %10.1f" % (n[:44], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(h, row))
        print("## %s  (id %s)" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        for i, n in enumerate(h):
            if n in KEYS:
                print("  %-84s %14s %s" % (n, row[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        rep(sys.argv[2])

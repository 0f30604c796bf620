#!/usr/bin/env python
"""per-source-line instruction and stall-sample totals of one kernel of an ncu report (needs -lineinfo + --import-source on)

  python profiles/srcpage.py <file.ncu-rep> <kernel-regex> [top]
"""
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
lines, h, fname = [], None, "?"
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif len(r) > 8 and r[0] == "Line No":
        h = r
        ci = {n: h.index(n) for n in ("Instructions Executed", "# Samples", "stall_long_sb", "L2 Theoretical Sectors Global")}
    elif h and len(r) == len(h) and r[0].isdigit():
        g = lambda n: int(r[ci[n]]) if r[ci[n]].isdigit() else 0
        lines.append((g("Instructions Executed"), g("# Samples"), g("stall_long_sb"), g("L2 Theoretical Sectors Global"),
                      "%s:%s" % (fname, r[0]), r[1].strip()))
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print("# %s  kernel %s: %d warp instructions, %d samples" % (rep, kern, tot_i, tot_s))
print("%7s %7s %7s %10s  %s" % ("inst%", "smpl%", "longsb%", "L2sectors", "line"))
for l in sorted(lines, key=lambda x: -x[0])[:top]:
    print("%6.2f%% %6.2f%% %6.2f%% %10d  %s: %s" % (100 * l[0] / tot_i, 100 * l[1] / tot_s, 100 * l[2] / tot_s, l[3], l[4], l[5][:105]))

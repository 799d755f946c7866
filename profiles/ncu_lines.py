#!/usr/bin/env python
"""Summarise an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`)
by CUDA source line: share of warp-stall samples, instructions executed, mean active threads."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and r[0] not in ("", "Function Name") and r[2] == "-":
        def g(name):
            try:
                return float(r[hdr[name]] or 0)
            except Exception:
                return 0.0
        lines.append((g("Warp Stall Sampling (All Samples)"), cur_file, r[0], r[1].strip()[:90], g("Instructions Executed"),
                      g("Avg. Threads Executed"), g("stall_long_sb"), g("stall_barrier"), g("stall_short_sb"), g("stall_wait")))
tot = sum(x[0] for x in lines) or 1
print("total samples %d" % tot)
for s, f, ln, src, ins, thr, lsb, bar, ssb, wt in sorted(lines, key=lambda x: -x[0])[:top]:
    print("%5.1f%% %-12s:%-4s inst=%-10d thr=%-4.1f long_sb=%-4.0f%% bar=%-3.0f%% | %s" %
          (100 * s / tot, f, ln, ins, thr, 100 * lsb / max(s, 1), 100 * bar / max(s, 1), src))

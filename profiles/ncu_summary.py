#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one block of key metrics per kernel launch,
and (with --traffic) a JSON object {kernel: dram bytes read+written per launch} for profiles/traffic.json."""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_requests_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    print("kernel:", name)
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            print("   %-85s %s %s" % (k, d[k], u.get(k, "")))
    def to_bytes(k):
        v = float(d[k].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[k]]
    try:
        traffic[name] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    except Exception as e:
        print("   (no dram bytes: %s)" % e)
if "--traffic" in sys.argv:
    print(json.dumps(traffic))

#!/usr/bin/env python
"""Prints the summary metrics, stall mix and hottest SASS lines of the first kernel in an .ncu-rep."""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
def page(p, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = page("raw"); h, r = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]
for n in want:
    if n in h: print("%-70s %s" % (n, r[h.index(n)]))
for i, n in enumerate(h):
    if "issue_stalled" in n and "per_issue_active" in n:
        v = float(r[i])
        if v >= 0.15: print("%-70s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), v))
rows = page("source", ["--print-source", "sass"])
hdr = rows[1]; data = []
for x in rows[2:]:
    if x and x[0] == "Kernel Name": break
    data.append(x)
ia, isamp, iex, ith = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
ilsb, iw, issb = hdr.index("stall_long_sb"), hdr.index("stall_wait"), hdr.index("stall_short_sb")
tot = sum(int(x[isamp]) for x in data)
print("total samples", tot, "instructions", len(data), "executed/warp", sum(int(x[iex]) for x in data) / max(1, int(data[0][iex])))
for k, x in enumerate(data):
    s = int(x[isamp])
    if s >= tot * thr:
        print("%4d %-64s samp %4d (%.1f%%) exec %7s thr %2s lsb %s ssb %s wait %s" % (k, x[ia].strip()[:64], s, 100.0 * s / tot, x[iex], x[ith], x[ilsb], x[issb], x[iw]))

#!/usr/bin/env python
"""Turns an .ncu-rep (captured on the B200 box with `ncu --set full --clock-control none
--import-source on`) into the text summary committed next to it.

    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r1_xyz_summary.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra),
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ["source: %s" % rep, ""]
    ki = hdr.index("Kernel Name")
    for r in data:
        lines.append("kernel: %s" % r[ki])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append("  %-66s %s %s" % (m, r[i], units[i]))
        rd, wr, t = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"),
                     hdr.index("gpu__time_duration.sum"))
        try:
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            b = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
            tt = float(r[t]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(units[t], 1.0)
            lines.append("  %-66s %.1f GB/s (traffic %.2f MB per launch)" % ("dram read+write / duration", b / tt / 1e9, b / 1e6))
        except ValueError:
            pass
        lines.append("")
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    print("wrote", dst)


if __name__ == "__main__":
    main()

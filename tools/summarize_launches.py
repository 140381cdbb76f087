#!/usr/bin/env python
"""Turns an ncu --csv launch list (several --metrics per launch) into one row per launch.
usage: python tools/summarize_launches.py gpurun_out/X_launches.csv > profiles/X_launches.txt"""
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    h = rows[0]
    iid, iname, igrid, imet, ival, iunit = (h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value", "Metric Unit"))
    by = {}
    order = []
    for r in rows[1:]:
        k = r[iid]
        if k not in by:
            by[k] = {"name": r[iname].split("(")[0].replace("void ", ""), "grid": r[igrid]}
            order.append(k)
        by[k][r[imet]] = (r[ival], r[iunit])
    print("%-4s %-28s %-14s %10s %12s %12s %14s %9s %7s %7s" % ("id", "kernel", "grid", "time_us", "dram_rd_MB", "dram_wr_MB", "warp_inst", "thr/inst", "L1hit%", "L2hit%"))
    tot = {}
    for k in order:
        d = by[k]

        def val(m, scale=1.0):
            if m not in d:
                return float("nan")
            v, u = d[m]
            v = float(v.replace(",", ""))
            mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            return v * mult * scale
        t = val("gpu__time_duration.sum")
        print("%-4s %-28s %-14s %10.1f %12.2f %12.2f %14.0f %9.2f %7.1f %7.1f" % (
            k, d["name"][:28], d["grid"], t, val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("sm__inst_executed.sum"),
            val("smsp__thread_inst_executed_per_inst_executed.ratio"), val("l1tex__t_sector_hit_rate.pct"), val("lts__t_sector_hit_rate.pct")))
        a = tot.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += val("dram__bytes_read.sum"); a[3] += val("dram__bytes_write.sum")
    all_t = sum(a[1] for a in tot.values())
    print()
    for n, a in tot.items():
        print("%-28s launches %3d  total %9.1f us  share %5.1f%%  dram rd+wr per launch %8.2f MB" % (n[:28], a[0], a[1], 100 * a[1] / all_t, (a[2] + a[3]) / a[0]))


if __name__ == "__main__":
    main()

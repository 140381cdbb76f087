#!/usr/bin/env python
"""ncu launch list of tools/ncu_workloads.py -> profiles/roofline_traffic.json (what bench.py quotes as roofline.traffic,
dram_gbs, issue_active_pct, threads_per_inst, hit rates; stamped with the hash of the kernel sources so that a stale file
is never quoted) and a readable table on stdout.
usage: python tools/make_roofline_traffic.py gpurun_out/TAG_workloads.csv gpurun_out/TAG_workloads.log > profiles/TAG_config_launches.txt"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    h = rows[0]
    iid, iname, imet, ival, iunit, igrid = (h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size"))
    launches, order = {}, []
    for r in rows[1:]:
        k = r[iid]
        if k not in launches:
            launches[k] = {"name": r[iname].split("(")[0].replace("void ", "").split("<")[0], "grid": int(r[igrid].strip("()").split(",")[0])}
            order.append(k)
        launches[k][r[imet]] = float(r[ival].replace(",", "")) * UNIT.get(r[iunit], 1.0)
    seqs, cur = [], None
    for k in order:                                         # a launch sequence starts with rr_prep_kernel
        L = launches[k]
        if L["name"] == "rr_prep_kernel":
            cur = []
            seqs.append(cur)
        if cur is not None:
            cur.append(L)
    labels = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")]
    out = {"kernel_source_hash": kernel_source_hash(), "source": os.path.basename(sys.argv[1]),
           "how": "ncu --clock-control none, tools/profile_configs.sh: second (warm) launch sequence of one 16-pose call per workload, one lane",
           "unit": "bytes per launch sequence of rr_trace_kernel (its per-pass launches summed: dram__bytes_read.sum + dram__bytes_write.sum)",
           "workloads": {}}
    # a call whose wave lists exceed the per-lane scratch budget runs as several launch sequences: a repetition is
    # complete when its draw launches cover the 16 x 400 items of the call
    reps, acc, items = [], [], 0
    for sq in seqs:
        acc += sq
        items += sum(L["grid"] for L in sq if L["name"] == "rr_draw_kernel")
        if items >= 16 * 400:
            reps.append(acc)
            acc, items = [], 0
    seqs = reps
    si = 0
    print("%-16s %-6s %9s %10s %10s %12s %8s %7s %7s %7s %8s %12s" % ("workload", "kernel", "time_us", "dram_rd_MB", "dram_wr_MB", "warp_inst", "thr/inst", "L1hit%", "L2hit%", "issue%", "warps%", "L2_wr_sect"))
    for lab in labels:
        seq = seqs[si + lab["sequences"] - 1]
        si += lab["sequences"]
        tr = [L for L in seq if L["name"] == "rr_trace_kernel"]
        dr = [L for L in seq if L["name"] == "rr_draw_kernel"]
        tsum = sum(L["gpu__time_duration.sum"] for L in tr)
        w = lambda key: sum(L[key] * L["gpu__time_duration.sum"] for L in tr) / tsum
        e = {"dram_bytes_per_launch": sum(L["dram__bytes_read.sum"] + L["dram__bytes_write.sum"] for L in tr),
             "trace_us": tsum, "draw_us": sum(L["gpu__time_duration.sum"] for L in dr),
             "issue_active_pct": w("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "threads_per_inst": [L["smsp__thread_inst_executed_per_inst_executed.ratio"] for L in tr],
             "l1_hit_pct": w("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": w("lts__t_sector_hit_rate.pct"),
             "warp_inst": sum(L["smsp__inst_executed.sum"] for L in tr),
             "local_ld_st_inst": sum(L["smsp__sass_inst_executed_op_local_ld.sum"] + L["smsp__sass_inst_executed_op_local_st.sum"] for L in tr),
             "casts": lab["casts"], "dram_bytes_per_cast": sum(L["dram__bytes_read.sum"] + L["dram__bytes_write.sum"] for L in tr) / max(lab["casts"], 1),
             "draw_l2_write_sectors": sum(L["lts__t_sectors_op_write.sum"] for L in dr),
             "draw_dram_bytes": sum(L["dram__bytes_read.sum"] + L["dram__bytes_write.sum"] for L in dr)}
        out["workloads"][lab["workload"]] = e
        for L in seq:
            if L["name"] in ("rr_trace_kernel", "rr_draw_kernel"):
                print("%-16s %-6s %9.1f %10.2f %10.2f %12.0f %8.2f %7.1f %7.1f %7.1f %8.1f %12.0f" % (
                    lab["workload"], L["name"][3:8], L["gpu__time_duration.sum"], L["dram__bytes_read.sum"] / 1e6, L["dram__bytes_write.sum"] / 1e6,
                    L["smsp__inst_executed.sum"], L["smsp__thread_inst_executed_per_inst_executed.ratio"], L["l1tex__t_sector_hit_rate.pct"],
                    L["lts__t_sector_hit_rate.pct"], L["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                    L["sm__warps_active.avg.pct_of_peak_sustained_active"], L["lts__t_sectors_op_write.sum"]))
        print("%-16s trace %.1f us, %.1f MB DRAM (%.1f B per cast), %.2f G casts/s under ncu; draw %.1f us" % (
            lab["workload"], tsum, e["dram_bytes_per_launch"] / 1e6, e["dram_bytes_per_cast"], lab["casts"] / tsum / 1e3, e["draw_us"]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` capture: headline metrics, stall reasons, hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep [top_n [launch_index]] > profiles/X_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, errors="replace").stdout


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which captured launch of the report (0-based)
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, units, r = rows[0], rows[1], rows[2 + which]
    print("# %s" % rep.split("/")[-1])
    print("kernel: %s" % r[h.index("Kernel Name")])
    for w in WANT:
        if w in h:
            i = h.index(w)
            print("%-66s %14s %s" % (w, r[i], units[i]))
    stalls = []
    for i, name in enumerate(h):
        if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
            try:
                stalls.append((float(r[i]), name.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    print("\nwarp stall reasons (pc samples):")
    for v, n in sorted(stalls, reverse=True)[:10]:
        print("  %5.1f%%  %s" % (100 * v / tot, n))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(which), "--launch-count", "1"]))))
    cur, hdr, ci, agg = None, None, None, {}
    for row in src:
        if len(row) == 2 and row[0] == "File Path":
            cur = row[1].split("/")[-1]
            continue
        if "# Samples" in row:
            hdr = row
            ci = (row.index("# Samples"), row.index("Instructions Executed"), row.index("Thread Instructions Executed"))
            continue
        if hdr is None or len(row) < len(hdr) or row[2] != "-":
            continue
        try:
            s, i, t = int(row[ci[0]] or 0), int(row[ci[1]] or 0), int(row[ci[2]] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur, row[0]), [0, 0, 0, row[1].strip()[:100]])
        a[0] += s; a[1] += i; a[2] += t
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print("\nhottest source lines (stall samples %, warp instructions %, active threads per instruction):")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("  %5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1), f, ln, a[3]))


if __name__ == "__main__":
    main()

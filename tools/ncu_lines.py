#!/usr/bin/env python
"""Aggregates an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python tools/ncu_lines.py src.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file, hdr, ci = None, None, None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if "# Samples" in r:
            hdr = r
            ci = {"samp": r.index("# Samples"), "inst": r.index("Instructions Executed"),
                  "tinst": r.index("Thread Instructions Executed")}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[2] != "-":          # SASS row (has an address): the CUDA row above already carries the per-line sums
            continue
        try:
            s, i, t = int(r[ci["samp"]] or 0), int(r[ci["inst"]] or 0), int(r[ci["tinst"]] or 0)
        except ValueError:
            continue
        key = (cur_file, r[0])
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:105]])
        a[0] += s; a[1] += i; a[2] += t
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print("total samples %d, warp instructions %d, files %s" % (ts, ti, sorted({k[0] for k in agg})))
    byfile = {}
    for (f, _), a in agg.items():
        b = byfile.setdefault(f, [0, 0]); b[0] += a[0]; b[1] += a[1]
    for f, b in byfile.items():
        print("  %-18s %5.1f%% samples %5.1f%% inst" % (f, 100.0 * b[0] / ts, 100.0 * b[1] / ti))
    print("--- top lines by stall samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("%5.1f%% samp %5.1f%% inst  %s:%s  %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, f, ln, a[3]))


if __name__ == "__main__":
    main()

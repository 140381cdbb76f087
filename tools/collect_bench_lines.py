#!/usr/bin/env python
"""gpurun_out/TAG_bench_*.json (one bench.py line each) -> profiles/TAG_bench_lines.jsonl + a short table on stdout."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = open(os.path.join(ROOT, "profiles", "%s_bench_lines.jsonl" % tag), "w")
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "%s_bench_*.json" % tag))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    d["_file"] = os.path.basename(f)
    out.write(json.dumps(d) + "\n")
    rows.append(d)
print("%-34s %5s %10s %10s %10s %9s %8s" % ("file", "gpus", "frames/s", "e2e", "ms/step", "single_ms", "scaling"))
for d in rows:
    print("%-34s %5d %10.1f %10.1f %10.3f %9s %8s" % (d["_file"][:34], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"],
                                                    ("%.3f" % d["single_frame_ms"]) if d.get("single_frame_ms") else "-", d["scaling"]))

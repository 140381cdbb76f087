set -x
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
RR_VERBOSE=1 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['config']['bvh_build_ms'], d['roofline']['nodes_visited'], d['roofline']['tris_tested'])
    else: print(line)"

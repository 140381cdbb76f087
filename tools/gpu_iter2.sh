cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --cpu-frames 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, d['e2e']['value'], d['roofline']['frac'])"

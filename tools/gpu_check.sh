cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --cpu-frames 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, 'e2e', d['e2e']['value'], 'trace_ms', r['kernel_ms'], 'draw_ms', r['draw_kernel_ms'], 'frac', r['frac'])"

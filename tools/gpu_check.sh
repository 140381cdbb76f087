# GPU box: the -m gpu suite with its full log kept, then one bench line. Usage: bash tools/gpu_check.sh [TAG]
cd /root/repo
TAG=${1:-check}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 --cpu-frames 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -1 gpurun_out/${TAG}_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, 'e2e', d['e2e']['value'], 'trace_ms', r['kernel_ms'], 'draw_ms', r['draw_kernel_ms'], 'frac', r['frac'])"

cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --steps 10 --warmup 3 --cpu-frames 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', {k:round(d[k],1) for k in ['value','ms_per_step']}, 'e2e', round(d['e2e']['value'],1), 'trace_ms', round(r['kernel_ms'],3), 'draw_ms', round(r['draw_kernel_ms'],3), 'frac', round(r['frac'],3))"; }
run base
for f in variants/lib_*.so; do RADARAYS_B200_LIB=$PWD/$f run $f; done

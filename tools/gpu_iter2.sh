cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" _mb8 _mb12; do
echo "== $v"
RADARAYS_B200_LIB=/root/repo/radarays_ros_b200/libradarays_b200$v.so python bench.py --steps 10 --warmup 3 --cpu-frames 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, d['e2e']['value'], d['roofline']['frac'])"
done

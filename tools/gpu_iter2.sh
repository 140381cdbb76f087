cd /root/repo
for v in mb2 mb4; do
echo "== $v"
RADARAYS_B200_LIB=/root/repo/radarays_ros_b200/libradarays_b200_$v.so python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','rays_bounces_per_s']}, d['e2e']['value'], d['roofline']['frac'])"
done
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rr_frame_kernel -s 17 -c 1 -o gpurun_out/prof_frame_v2 python bench.py --steps 2 --warmup 1 > gpurun_out/prof_v2.log 2>&1

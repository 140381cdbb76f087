# 1 GPU: one bench line per BASELINE config (2 default with the CPU baseline, 3 at 2048 and 256 samples, 4 both shardings, 5).
cd /root/repo; TAG=${1:-r2}
python bench.py                                              > gpurun_out/${TAG}_bench_cfg2_n1.json 2> gpurun_out/${TAG}_bench_cfg2_n1.err
python bench.py --config 3 --cpu-frames 0                    > gpurun_out/${TAG}_bench_cfg3_s2048_n1.json 2> gpurun_out/${TAG}_bench_cfg3_s2048_n1.err
python bench.py --config 3 --samples 64 --cpu-frames 0       > gpurun_out/${TAG}_bench_cfg3_s64_n1.json 2> gpurun_out/${TAG}_bench_cfg3_s64_n1.err
python bench.py --config 4 --shard pose --cpu-frames 0       > gpurun_out/${TAG}_bench_cfg4_pose_n1.json 2> gpurun_out/${TAG}_bench_cfg4_pose_n1.err
python bench.py --config 4 --shard azimuth --cpu-frames 0    > gpurun_out/${TAG}_bench_cfg4_az_p2p_n1.json 2> gpurun_out/${TAG}_bench_cfg4_az_p2p_n1.err
python bench.py --config 5 --steps 2 --warmup 1 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg5_n1.json 2> gpurun_out/${TAG}_bench_cfg5_n1.err
python bench.py --impl reference --steps 6 --warmup 1        > gpurun_out/${TAG}_bench_reference_n1.json 2> gpurun_out/${TAG}_bench_reference_n1.err
for f in gpurun_out/${TAG}_bench_*_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline',{})
    print(sys.argv[1].split('/')[-1], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'pageable', d['e2e'].get('pageable_value'), 'single', d.get('single_frame_ms'), 'trace', r.get('kernel_ms'), 'draw', r.get('draw_kernel_ms'), 'frac', r.get('frac'), 'dram_frac', r.get('dram_frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'bvh_ms', d.get('config',{}).get('bvh_build_ms'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done

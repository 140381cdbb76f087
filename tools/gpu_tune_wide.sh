# GPU box: the 4-wide BVH build variant (RR_WIDE_BVH=1, variants/lib_wide*.so): parity tests through the variant library,
# then a short bench of every variant next to the default binary-BVH library. Usage: gpurun -- 'bash tools/gpu_tune_wide.sh [steps]'
cd /root/repo
STEPS=${1:-100}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  RADARAYS_B200_LIB=$PWD/variants/lib_wide.so python -m pytest tests/test_ray_triangle_kat.py tests/test_gpu_parity.py tests/test_gpu_wavefront.py tests/test_gpu_fullsize.py tests/test_golden_frames.py -m gpu -q -x --tb=short > gpurun_out/wide_pytest.log 2>&1
  tail -4 gpurun_out/wide_pytest.log
fi
run() {
  python bench.py --steps $STEPS --warmup 3 --cpu-frames 0 2> gpurun_out/tune_$1.err | tail -1 > gpurun_out/tune_$1.json
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/tune_%s.json' % tag).read()); r = d['roofline']; e = d['e2e']
    print('%-10s value %7.0f ms %.3f | e2e %7.0f | trace %.3f draw %.3f | nodes/cast %.1f tris/cast %.2f node_bytes %s bvh_ms %.0f | sum %d' % (
        tag, d['value'], d['ms_per_step'], e['value'], r['kernel_ms'], r['draw_kernel_ms'], r['nodes_per_cast'], r['tris_per_cast'],
        r.get('node_bytes'), d['config']['bvh_build_ms'], d['image_checksum']))
except Exception as ex:
    print(tag, 'FAILED', ex)
PY
}
run base
for f in variants/lib_*.so; do t=$(basename $f .so); RADARAYS_B200_LIB=$PWD/$f run ${t#lib_}; done

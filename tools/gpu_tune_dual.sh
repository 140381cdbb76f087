# GPU box: two rays per lane (RR_PASS_DUAL=1: rr_dual_kernel) against the fused rr_trace_kernel: parity tests under the
# switch, then one short bench per occupancy variant. Usage: gpurun -- 'bash tools/gpu_tune_dual.sh [steps]'
cd /root/repo
STEPS=${1:-100}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  RR_PASS_DUAL=1 python -m pytest tests/test_ray_triangle_kat.py tests/test_gpu_parity.py tests/test_gpu_wavefront.py tests/test_gpu_fullsize.py tests/test_golden_frames.py tests/test_gpu_gen_radar_image.py -m gpu -q -x --tb=short > gpurun_out/dual_pytest.log 2>&1
  tail -4 gpurun_out/dual_pytest.log
fi
run() {
  python bench.py --steps $STEPS --warmup 3 --cpu-frames 0 2> gpurun_out/tune_$1.err | tail -1 > gpurun_out/tune_$1.json
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/tune_%s.json' % tag).read()); r = d['roofline']; e = d['e2e']
    print('%-10s value %7.0f ms %.3f | e2e %7.0f call16 %7.0f | trace %.3f draw %.3f | nodes/cast %.1f | sum %d' % (
        tag, d['value'], d['ms_per_step'], e['value'], e.get('call16_value', 0), r['kernel_ms'], r['draw_kernel_ms'], r['nodes_per_cast'], d['image_checksum']))
except Exception as ex:
    print(tag, 'FAILED', ex)
PY
}
RR_PASS_DUAL=0 run fused
RR_PASS_DUAL=1 run dual6
for f in variants/lib_d*.so; do t=$(basename $f .so); RR_PASS_DUAL=1 RADARAYS_B200_LIB=$PWD/$f run dual_${t#lib_}; done

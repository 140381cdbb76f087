# GPU box: the two-kernel form of a pass (RR_PASS_SPLIT=1: rr_walk_kernel + rr_shade_kernel) against the fused rr_trace_kernel:
# the -m gpu suite under the split, then one short bench per occupancy variant. Usage: gpurun -- 'bash tools/gpu_tune_split.sh [steps]'
cd /root/repo
STEPS=${1:-100}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then RR_PASS_SPLIT=1 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/split_pytest.log 2>&1; tail -4 gpurun_out/split_pytest.log; fi
run() {
  python bench.py --steps $STEPS --warmup 3 --cpu-frames 0 2> gpurun_out/tune_$1.err | tail -1 > gpurun_out/tune_$1.json
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/tune_%s.json' % tag).read()); r = d['roofline']; e = d['e2e']
    print('%-10s value %7.0f ms %.3f | e2e %7.0f call16 %7.0f pageable %7.0f | trace %.3f draw %.3f | launches %d sum %d' % (
        tag, d['value'], d['ms_per_step'], e['value'], e.get('call16_value', 0), e.get('pageable_value', 0),
        r['kernel_ms'], r['draw_kernel_ms'], d['gpu_launches'], d['image_checksum']))
except Exception as ex:
    print(tag, 'FAILED', ex)
PY
}
RR_PASS_SPLIT=0 run fused
RR_PASS_SPLIT=1 run split_w12s8
for f in variants/lib_*.so; do t=$(basename $f .so); RR_PASS_SPLIT=1 RADARAYS_B200_LIB=$PWD/$f run split_${t#lib_}; done

# GPU box: full -m gpu suite with the default library, then one short bench per tuning variant (variants/lib_*.so,
# built here by tools/build_variants.sh or by hand). Usage: gpurun -- 'bash tools/gpu_tune2.sh [steps]'
cd /root/repo
STEPS=${1:-100}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/tune_pytest.log 2>&1; fi
tail -5 gpurun_out/tune_pytest.log
run() {
  python bench.py --steps $STEPS --warmup 3 --cpu-frames 0 2> gpurun_out/tune_$1.err | tail -1 > gpurun_out/tune_$1.json
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/tune_%s.json' % tag).read()); r = d['roofline']; e = d['e2e']
    print('%-10s value %7.0f ms %.3f | e2e %7.0f call16 %7.0f pageable %7.0f | trace %.3f draw %.3f | sum %d' % (
        tag, d['value'], d['ms_per_step'], e['value'], e.get('call16_value', 0), e.get('pageable_value', 0),
        r['kernel_ms'], r['draw_kernel_ms'], d['image_checksum']))
except Exception as ex:
    print(tag, 'FAILED', ex)
PY
}
run base
for f in variants/lib_*.so; do t=$(basename $f .so); RADARAYS_B200_LIB=$PWD/$f run ${t#lib_}; done

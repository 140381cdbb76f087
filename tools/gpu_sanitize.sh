# compute-sanitizer over the small-scene GPU tests (all three kernels, ragged wave lists, chunked draw lists, GenRadarImage
# goals, sharded column-major output). Logs -> gpurun_out/TAG_sanitizer_*.log (copy to profiles/). Usage: bash tools/gpu_sanitize.sh TAG
cd /root/repo; TAG=${1:-r2}
T="tests/test_gpu_wavefront.py tests/test_gpu_gen_radar_image.py tests/test_gpu_api_errors.py tests/test_golden_frames.py"
S="tests/test_gpu_param_sweep.py -k 0-or-1-or-2-or-3-or-4-or-5"
for tool in memcheck racecheck initcheck; do
  extra=""; [ $tool = initcheck ] && extra=""
  timeout 1500 compute-sanitizer --tool $tool $extra --log-file gpurun_out/${TAG}_sanitizer_${tool}.log --print-limit 20 \
      python -m pytest $T -m gpu -q -x --tb=line -p no:cacheprovider > gpurun_out/${TAG}_sanitizer_${tool}_pytest.log 2>&1
  echo "== $tool: $(tail -1 gpurun_out/${TAG}_sanitizer_${tool}_pytest.log)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -2
done
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_sanitizer_memcheck_sweep.log --print-limit 20 \
    python -m pytest tests/test_gpu_param_sweep.py -m gpu -q -x --tb=line -p no:cacheprovider > gpurun_out/${TAG}_sanitizer_memcheck_sweep_pytest.log 2>&1
echo "== memcheck sweep: $(tail -1 gpurun_out/${TAG}_sanitizer_memcheck_sweep_pytest.log)"; grep -E "ERROR SUMMARY" gpurun_out/${TAG}_sanitizer_memcheck_sweep.log | tail -1

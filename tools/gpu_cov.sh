cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
python tools/run_configs.py > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs_r1.err
cat gpurun_out/configs_r1.jsonl; tail -3 gpurun_out/configs_r1.err

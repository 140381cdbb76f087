# full-set ncu captures of the three rr_trace_kernel passes of one launch sequence. Usage: bash tools/gpu_trace_profile.sh TAG
cd /root/repo; TAG=${1:-trace}
ncu --set full --clock-control none --import-source on -k regex:"rr_trace_kernel" --launch-skip 9 -c 3 -f -o gpurun_out/${TAG} \
    python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/${TAG}.log 2>&1
ls -la gpurun_out/${TAG}*

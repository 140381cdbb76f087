# one full-set ncu capture of rr_draw_kernel inside a short bench run. Usage: bash tools/gpu_draw_profile.sh TAG
cd /root/repo; TAG=${1:-draw}
ncu --set full --clock-control none --import-source on -k regex:"rr_draw_kernel" --launch-skip 3 -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/${TAG}.log 2>&1
ls -la gpurun_out/${TAG}*

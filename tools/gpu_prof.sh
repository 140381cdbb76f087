cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"rr_draw_kernel" --launch-skip 17 -c 1 -f -o gpurun_out/draw_v8 python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/draw_v8.log 2>&1

cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rr_draw_kernel -s 17 -c 1 -o gpurun_out/prof_draw_v5 python bench.py --steps 2 --warmup 1 --cpu-frames 0 > gpurun_out/prof_v5d.log 2>&1

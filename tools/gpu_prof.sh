cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"rr_trace_kernel|rr_scan_kernel|rr_draw_kernel" --launch-skip 104 -c 1 -o gpurun_out/prof_trace_v6_p1 -f python bench.py --steps 2 --warmup 1 --cpu-frames 0 > gpurun_out/prof_v6.log 2>&1
tail -3 gpurun_out/prof_v6.log | cut -c1-300

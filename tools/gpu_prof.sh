cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 3 --warmup 1 --cpu-frames 0 > gpurun_out/l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rr_trace_kernel -s 17 -c 1 -o gpurun_out/prof_trace_v4 python bench.py --steps 2 --warmup 1 --cpu-frames 0 > gpurun_out/prof_v4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rr_draw_kernel -s 17 -c 1 -o gpurun_out/prof_draw_v4 python bench.py --steps 2 --warmup 1 --cpu-frames 0 > gpurun_out/prof_v4d.log 2>&1
ls -la gpurun_out | tail -5

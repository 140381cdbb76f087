cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rr_frame_kernel -s 17 -c 1 -o gpurun_out/prof_frame_v3 python bench.py --steps 2 --warmup 1 > gpurun_out/prof_v3.log 2>&1
ls -la gpurun_out | tail -3

set -x
cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rr_frame_kernel -s 17 -c 1 -o gpurun_out/prof_frame python bench.py --steps 2 --warmup 1 > gpurun_out/prof.log 2>&1
ls -la gpurun_out
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1_a.json
cat gpurun_out/bench_r1_a.json

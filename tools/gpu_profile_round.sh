# Round profile capture: (1) launch list of one bench command with per-launch duration / DRAM bytes / instruction counts,
# (2) one full-set capture each of the pass-1 trace launch and of the draw launch. Usage: bash tools/gpu_profile_round.sh TAG
cd /root/repo
TAG=${1:-r1}
mkdir -p gpurun_out
# the bench first runs 16 single-pose stats frames (6 launches each) and the BVH build; skip them
SKIP=${SKIP:-102}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:"rr_trace_kernel|rr_scan_kernel|rr_draw_kernel" --launch-skip $SKIP -c 24 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rr_trace_kernel" --launch-skip $((SKIP/2 + 1)) -c 1 -f -o gpurun_out/${TAG}_trace_pass1 \
    python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/${TAG}_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rr_draw_kernel" --launch-skip 17 -c 1 -f -o gpurun_out/${TAG}_draw \
    python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/${TAG}_draw.log 2>&1
ls -la gpurun_out | grep ${TAG}

cd /root/repo
mkdir -p gpurun_out
for v in r32 r8w10; do
RADARAYS_B200_LIB=$PWD/variants/lib_$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:"rr_walk_kernel|rr_shade_kernel|rr_scan_kernel|rr_draw_kernel" --launch-skip 153 -c 9 --csv \
    --log-file gpurun_out/v7_${v}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-frames 0 --lanes 1 > gpurun_out/v7_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/v7_${v}_launches.csv
done

# ncu launch list (durations, DRAM bytes, hit rates, issue utilisation, divergence, L2 write sectors) of one launch sequence of
# BASELINE configs 2, 3 (64..2048 samples) and 4. Usage: bash tools/profile_configs.sh TAG ; then
#   python tools/make_roofline_traffic.py gpurun_out/TAG_workloads.csv gpurun_out/TAG_workloads.log
cd /root/repo; TAG=${1:-r2}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_write.sum
M=$M,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"rr_prep_kernel|rr_trace_kernel|rr_scan_kernel|rr_draw_kernel" --csv \
    --log-file gpurun_out/${TAG}_workloads.csv python tools/ncu_workloads.py $2 > gpurun_out/${TAG}_workloads.log 2>&1
tail -3 gpurun_out/${TAG}_workloads.log; wc -l gpurun_out/${TAG}_workloads.csv

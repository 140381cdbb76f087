# ncu launch lists (one warm 16-pose sequence of config 2) of the measured alternatives to the fused binary walk:
# the 4-wide BVH build (variants/lib_wide.so), two rays per lane (RR_PASS_DUAL=1), cast / shading split (RR_PASS_SPLIT=1).
# Usage: gpurun -- 'bash tools/gpu_alt_launches.sh TAG'  ->  gpurun_out/TAG_alt_<name>.csv
cd /root/repo; TAG=${1:-r2b}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,launch__registers_per_thread
K='regex:rr_trace_kernel|rr_walk_kernel|rr_shade_kernel|rr_dual_kernel'
run() { name=$1; shift; env "$@" ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/${TAG}_alt_$name.csv python tools/ncu_workloads.py --config2-only > gpurun_out/${TAG}_alt_$name.log 2>&1; tail -1 gpurun_out/${TAG}_alt_$name.log; }
run fused RR_PASS_DUAL=0
run wide RADARAYS_B200_LIB=$PWD/variants/lib_wide.so
run dual RR_PASS_DUAL=1
run split RR_PASS_SPLIT=1

# Final evidence of a round, one GPU: launch lists of all configs (-> roofline_traffic.json), full ncu captures of the trace
# passes and the draw kernel, then one bench line per config (which now quote the fresh, hash-stamped ncu numbers).
cd /root/repo; TAG=${1:-r2}
bash tools/profile_configs.sh $TAG
python tools/make_roofline_traffic.py gpurun_out/${TAG}_workloads.csv gpurun_out/${TAG}_workloads.log > gpurun_out/${TAG}_config_launches.txt
cp profiles/roofline_traffic.json gpurun_out/${TAG}_roofline_traffic.json
bash tools/gpu_trace_profile.sh ${TAG}_trace_final
bash tools/gpu_draw_profile.sh ${TAG}_draw_final
bash tools/gpu_single.sh $TAG

# builds tuning variants of the kernels into gpurun_scratch/lib_<tag>.so (travels to the GPU box; git-ignored)
# usage: bash tools/build_variants.sh "tag1:-DRR_REFILL=4" "tag2:-DRR_REFILL=16 -DRR_WALK_MIN_BLOCKS=10"
set -e
cd /root/repo/radarays_ros_b200/csrc
mkdir -p ../../variants
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-mfma,-O2"
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  $NV $flags -c rr_kernels.cu -o /tmp/rr_kernels_$tag.o
  $NV -shared -o ../../variants/lib_$tag.so rr_api.o /tmp/rr_kernels_$tag.o rr_bvh_build.o rr_bvh.o rr_mesh_io.o -lcudart_static -lpthread -ldl -lrt
  echo built variants/lib_$tag.so "($flags)"
done

# like build_variants.sh but rebuilds every object (for macros that touch the BVH builders too)
# usage: bash tools/build_variants_full.sh "tag:-DRR_MAX_LEAF=8" ...
set -e
cd /root/repo/radarays_ros_b200/csrc
mkdir -p ../../variants
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-mfma,-O2"
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  for f in rr_api rr_kernels rr_bvh_build; do $NV $flags -c $f.cu -o /tmp/${f}_$tag.o & done
  $NV $flags -x cu -c rr_bvh.cpp -o /tmp/rr_bvh_$tag.o &
  wait
  $NV -shared -o ../../variants/lib_$tag.so /tmp/rr_api_$tag.o /tmp/rr_kernels_$tag.o /tmp/rr_bvh_build_$tag.o /tmp/rr_bvh_$tag.o rr_mesh_io.o -lcudart_static -lpthread -ldl -lrt
  echo built variants/lib_$tag.so "($flags)"
done

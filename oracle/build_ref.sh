#!/usr/bin/env bash
# Builds oracle/_ref/libradarays_ref.so from the reference's OWN sources where they lie (/root/reference), compiled
# unmodified against oracle/ref_shim (see its README). Outputs only into oracle/_ref/ (git-ignored, shipped by gpurun).
# Only runs where /root/reference exists (this container); the GPU box uses the prebuilt .so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RADARAYS_REFERENCE:-/root/reference}"
[ -d "$REF/src/radarays_ros" ] || { echo "build_ref.sh: $REF not present, skipping"; exit 0; }
mkdir -p "$HERE/_ref"
CXX="${HOSTCXX:-/usr/bin/g++}"
FLAGS="-O3 -march=x86-64-v3 -std=c++17 -fPIC -fopenmp -ffp-contract=off -mfma -w -include $HERE/ref_shim/pre.h -I$HERE/ref_shim -I$REF/include"
for f in RadarCPU Radar radar_algorithms; do
  $CXX $FLAGS -c "$REF/src/radarays_ros/$f.cpp" -o "$HERE/_ref/$f.o"
done
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$HERE/_ref/ref_harness.o"
$CXX -shared -fopenmp -o "$HERE/_ref/libradarays_ref.so" "$HERE/_ref/RadarCPU.o" "$HERE/_ref/Radar.o" "$HERE/_ref/radar_algorithms.o" "$HERE/_ref/ref_harness.o"
echo "built $HERE/_ref/libradarays_ref.so"
# the reference-side binding (radarays_ros_b200/cpp/RadarB200.hpp) against the reference's own Radar base class
B200="$HERE/../radarays_ros_b200"
if [ -f "$B200/libradarays_b200.so" ]; then
  $CXX $FLAGS -I"$REF/include/radarays_ros" -I"$HERE/../include" -c "$HERE/adapter_harness.cpp" -o "$HERE/_ref/adapter_harness.o"
  $CXX -shared -fopenmp -o "$HERE/_ref/libradarays_adapter.so" "$HERE/_ref/Radar.o" "$HERE/_ref/adapter_harness.o" \
      -L"$B200" -lradarays_b200 -Wl,-rpath,'$ORIGIN/../../radarays_ros_b200'
  echo "built $HERE/_ref/libradarays_adapter.so"
fi
rm -f "$HERE"/_ref/*.o

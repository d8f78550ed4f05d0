#!/bin/bash
# Build kernel variants side by side (compile-time -D switches added for an experiment) — run here (nvcc, no GPU):
#   benchmarks/variants.sh name:"-DFLAG=1" ...   -> adsurf_b200/_lib/variants/lib_<name>.so
# and measure them on the GPU box with benchmarks/run_variants.sh (each is loaded through ADSURF_LIB).
set -e
cd "$(dirname "$0")/.."
mkdir -p adsurf_b200/_lib/variants
build() {
  name=$1; shift
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v "$@" \
       -I include -I adsurf_b200/csrc -o adsurf_b200/_lib/variants/lib_$name.so adsurf_b200/csrc/adsurf_kernels.cu \
       > adsurf_b200/_lib/variants/build_$name.log 2>&1
  grep -A2 "k_grid_bwd" adsurf_b200/_lib/variants/build_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
}
# usage: benchmarks/variants.sh name1:"-DFLAG=1 ..." name2:"..."   (the current sources define no switches;
# profiles/r2_variants.txt records what the round-2 switches measured before they were removed)
build base &
for spec in "$@"; do
  build "${spec%%:*}" ${spec#*:} &
done
wait

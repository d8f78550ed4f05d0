#!/bin/bash
# Build kernel variants side by side (compile-time switches of the dense adjoint sweep) — run here (nvcc, no GPU):
#   benchmarks/variants.sh            -> adsurf_b200/_lib/variants/lib_<name>.so
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
build base &
build noclobber -DADSURF_VARIANT_NOCLOBBER=1 &
build noearly -DADSURF_VARIANT_FETCH_E_EARLY=0 &
build late -DADSURF_VARIANT_SINK_LATE=1 &
wait

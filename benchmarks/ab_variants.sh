#!/bin/bash
# A/B of variant libraries in one GPU session, interleaved: benchmarks/ab_variants.sh <models> <rounds> name1 name2 ...
mkdir -p gpurun_out; : > gpurun_out/ab.txt
M=$1; R=$2; shift; shift
for r in $(seq 1 $R); do
  for v in "$@"; do
    echo -n "$v " >> gpurun_out/ab.txt
    ADSURF_LIB=$PWD/adsurf_b200/_lib/variants/lib_$v.so python benchmarks/quick_line.py $M 6 >> gpurun_out/ab.txt 2>&1
  done
done
cat gpurun_out/ab.txt

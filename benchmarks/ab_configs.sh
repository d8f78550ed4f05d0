#!/bin/bash
# A/B of variant libraries on the end-to-end configs (interleaved): benchmarks/ab_configs.sh <rounds> name1 name2 ...
mkdir -p gpurun_out; : > gpurun_out/ab_configs.txt
R=$1; shift
for r in $(seq 1 $R); do
  for v in "$@"; do
    ADSURF_LIB=$PWD/adsurf_b200/_lib/variants/lib_$v.so python benchmarks/configs_e2e.py 2>/dev/null | \
      python -c "
import sys, json
out = ['$v']
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); out.append('c%d %.0f' % (d['config'], d['it_per_s']))
print(' '.join(out))" >> gpurun_out/ab_configs.txt
  done
done
cat gpurun_out/ab_configs.txt

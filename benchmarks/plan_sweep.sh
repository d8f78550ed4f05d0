#!/bin/bash
# Forward-kernel plans tried by hand on one config (ADSURF_PLAN_GRID_FWD = warps per block, chunk, velocities per fill):
#   benchmarks/plan_sweep.sh <config: 2|3> "w,chunk,tc" ...
mkdir -p gpurun_out; : > gpurun_out/plan_sweep.txt
CFG=$1; shift
for p in default "$@"; do
  if [ "$p" = default ]; then unset ADSURF_PLAN_GRID_FWD; else export ADSURF_PLAN_GRID_FWD=$p; fi
  python benchmarks/configs_e2e.py --only $CFG 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('$p', 'config', d['config'], round(d['it_per_s']), 'it/s')" >> gpurun_out/plan_sweep.txt
done
cat gpurun_out/plan_sweep.txt

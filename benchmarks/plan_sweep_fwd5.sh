#!/bin/bash
# Forward-only dense sweep of a config-5 shard under forced plans: benchmarks/plan_sweep_fwd5.sh <models> "w,chunk,tc" ...
mkdir -p gpurun_out; : > gpurun_out/plan_sweep_fwd5.txt
M=$1; shift
for p in default "$@"; do
  if [ "$p" = default ]; then unset ADSURF_PLAN_GRID_FWD; else export ADSURF_PLAN_GRID_FWD=$p; fi
  python bench.py --models $M --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$M models', '$p', round(d['forward_only']['value']/1e9,4), 'G samples/s')" >> gpurun_out/plan_sweep_fwd5.txt
done
cat gpurun_out/plan_sweep_fwd5.txt

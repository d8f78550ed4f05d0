#!/bin/bash
# On the GPU box: parity subset + headline sweep for every variant library (see benchmarks/variants.sh)
mkdir -p gpurun_out
: > gpurun_out/variants.log
for lib in adsurf_b200/_lib/variants/lib_*.so; do
  name=$(basename $lib .so)
  ADSURF_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "grid_golden or config5_subset or ragged" 2>&1 | tail -1 >> gpurun_out/variants.log
  ADSURF_LIB=$PWD/$lib timeout 600 python bench.py --models ${VMODELS:-1024} --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$name', round(j['value']/1e9,4), 'G/s fwd+bwd', round(j['forward_only']['value']/1e9,3), 'G/s fwd', j['clocks'])" >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log

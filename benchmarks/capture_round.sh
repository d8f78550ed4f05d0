#!/bin/bash
# Round artefacts in one GPU call (run from the repo root on the GPU box, e.g. through gpurun):
#   benchmarks/capture_round.sh [tests] [bench] [configs] [launches] ...
# Outputs go to gpurun_out/; copy what should be judged into profiles/.
# ncu runs only after the same command line exited 0 without it (B200_PROFILING.md).
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2>&1 ;;
configs)
  timeout 900 python benchmarks/configs_e2e.py > gpurun_out/configs_e2e.jsonl 2> gpurun_out/configs_e2e.err ;;
launches)
  timeout 600 python benchmarks/configs_e2e.py --iters2 24 --iters3 24 --iters4 16 > gpurun_out/configs_short.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_configs.csv \
      python benchmarks/configs_e2e.py --iters2 24 --iters3 24 --iters4 16 > gpurun_out/ncu_launch_configs.log 2>&1 ;;
launches_bench)
  timeout 600 python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_512.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench512.csv \
      python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_bench.log 2>&1 ;;
esac
done

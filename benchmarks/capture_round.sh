#!/bin/bash
# Round artefacts in one GPU call (run from the repo root on the GPU box, e.g. through gpurun):
# GPU tests, smoke, the full N=1 bench, the reference arm, the per-launch list of the bench command (ncu, only after
# the same command exited 0 without it) and the end-to-end iteration rates of BASELINE configs 2-4.
# Outputs go to gpurun_out/; copy what should be judged into profiles/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_r1.log 2>gpurun_out/bench_full_r1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.log 2>&1
python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_512.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python benchmarks/configs_e2e.py > gpurun_out/configs_e2e_r1.log 2>&1

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
quick)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err ;;
ncu_bwd)
  timeout 600 python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_512.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_grid_(bwd|fwd)' -s 4 -c 2 -f -o gpurun_out/prof_r2_grid \
      python bench.py --models 512 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_grid.log 2>&1 ;;
ncu_points)
  timeout 600 python benchmarks/configs_e2e.py --iters2 24 --iters3 24 --iters4 16 > gpurun_out/configs_short.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_points_bwd|k_ens_|k_grid_fwd' -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-12} -f -o gpurun_out/prof_r2_points \
      python benchmarks/configs_e2e.py --iters2 24 --iters3 24 --iters4 16 > gpurun_out/ncu_points.log 2>&1 ;;
ncu_c4)
  # pointwise DMF kernel of config 4 (1024 and 4096 models): with --no-warm the launches are the InversionLoop's
  # warm step (2 kernels) + iterations x 2; skip the first 6 matching kernels, capture 2 iterations
  for M in 1024 4096; do
    timeout 600 python benchmarks/configs_e2e.py --only 4 --models4 $M --iters4 8 --no-warm > gpurun_out/c4_$M.log 2>&1 && \
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_points_bwd|k_ens_' -s 6 -c 4 -f \
        -o gpurun_out/prof_r2_config4_$M python benchmarks/configs_e2e.py --only 4 --models4 $M --iters4 8 --no-warm \
        > gpurun_out/ncu_c4_$M.log 2>&1
  done ;;
ncu_c3)
  timeout 600 python benchmarks/configs_e2e.py --only 3 --iters3 8 --no-warm > gpurun_out/c3.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_points_bwd|k_ens_|k_grid_fwd' -s 9 -c 6 -f \
      -o gpurun_out/prof_r2_config3 python benchmarks/configs_e2e.py --only 3 --iters3 8 --no-warm > gpurun_out/ncu_c3.log 2>&1 ;;
ncu_full)
  # one full-size launch of the headline kernel (DRAM traffic of the launch the bench times)
  timeout 600 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_1step.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_grid_bwd' -s 3 -c 1 -f -o gpurun_out/prof_r2_full \
      python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ;;
persist)
  python benchmarks/quick_line.py 1024 4 > gpurun_out/persist_l2.txt 2>&1
  ADSURF_PERSIST_L2=0 python benchmarks/quick_line.py 1024 4 >> gpurun_out/persist_l2.txt 2>&1
  python benchmarks/quick_line.py 1024 4 det >> gpurun_out/persist_l2.txt 2>&1
  ADSURF_PERSIST_L2=0 python benchmarks/quick_line.py 1024 4 det >> gpurun_out/persist_l2.txt 2>&1 ;;
configs_persist)
  timeout 900 python benchmarks/configs_e2e.py --only 4 > gpurun_out/configs_c4_nopersist.jsonl 2>&1
  ADSURF_PERSIST_L2=1 timeout 900 python benchmarks/configs_e2e.py --only 4 > gpurun_out/configs_c4_persist.jsonl 2>&1 ;;
dispersion)
  timeout 900 python benchmarks/dispersion_bench.py > gpurun_out/dispersion_bench.json 2> gpurun_out/dispersion_bench.err ;;
ncu_disp)
  timeout 300 python benchmarks/dispersion_bench.py --profile ${DISP_MODELS:-592} > gpurun_out/disp_profile.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_dispersion' -s 1 -c 1 -f -o gpurun_out/prof_r2_dispersion \
      python benchmarks/dispersion_bench.py --profile ${DISP_MODELS:-592} > gpurun_out/ncu_disp.log 2>&1 ;;
variants)
  bash benchmarks/run_variants.sh > gpurun_out/run_variants.out 2>&1 ;;
sharded)
  N=${NGPU:-2}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      benchmarks/check_sharded.py > gpurun_out/check_sharded_n$N.json 2> gpurun_out/check_sharded_n$N.err; echo "rc=$?" >> gpurun_out/check_sharded_n$N.err ;;
multi)
  N=${NGPU:-2}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err ;;
esac
done

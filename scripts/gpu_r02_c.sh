#!/bin/bash
# round 2, call C (1 GPU): GPU suite after the stack fixes; the per-rank share of the 8-GPU strong-scaling step
# (13 poses) measured on one GPU, with its launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --batch 13 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_b13.log 2>&1; tail -c 1200 gpurun_out/r02_bench_c2_b13.log
timeout 600 python bench.py --batch 25 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_b25.log 2>&1; tail -c 600 gpurun_out/r02_bench_c2_b25.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_b13.csv \
    python bench.py --batch 13 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_b13.log 2>&1
tail -30 gpurun_out/r02_launches_b13.csv | cut -c1-200

#!/bin/bash
# One gpurun call while iterating on kernels: pipe ubench (if built), GPU tests, layout sweep, C2 bench.
set -x
mkdir -p gpurun_out
[ -x scripts/ubench/pipes ] && timeout 120 scripts/ubench/pipes > gpurun_out/pipes.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python scripts/sweep_layouts.py > gpurun_out/sweep.log 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
tail -2 gpurun_out/bench_c2.log

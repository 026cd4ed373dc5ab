#!/bin/bash
# One gpurun call while iterating: selected GPU tests ($1 = pytest -k expression or empty), C2 bench without the CPU leg.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
tail -2 gpurun_out/bench_c2.log

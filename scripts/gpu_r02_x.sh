#!/bin/bash
# round 2, call X (1 GPU): nearest-neighbour interpolation on the GPU + the DRR / pipeline suites
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drr.py tests/test_gpu_pipeline.py tests/test_gpu_adapters_run.py -m gpu -x -q > gpurun_out/pytest_x.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_x.log

#!/bin/bash
# round 2, call AG (1 GPU): log remap on the device against the oracle
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_preproc.py -m gpu -x -q > gpurun_out/pytest_ag.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_ag.log

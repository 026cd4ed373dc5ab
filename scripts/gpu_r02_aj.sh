#!/bin/bash
# round 2, call AJ (1 GPU): golden vectors of the round-2 additions on the device
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_aj.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_aj.log

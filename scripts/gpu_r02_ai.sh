#!/bin/bash
# round 2, call AI (1 GPU): the C++ host mirror test with the depth ray caster and the pre-processing
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cpp_mirror.py -m gpu -x -q -s > gpurun_out/pytest_ai.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_ai.log

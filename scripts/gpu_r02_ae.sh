#!/bin/bash
# round 2, call AE (1 GPU): scaled-up randomised sweep (3x the cases) with its pass count
set -x
mkdir -p gpurun_out
XREG_FUZZ_SCALE=3 timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q > gpurun_out/r02_fuzz_scale3.log 2>&1; echo "fuzz exit $?" >> gpurun_out/r02_fuzz_scale3.log
tail -4 gpurun_out/r02_fuzz_scale3.log

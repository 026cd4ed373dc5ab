#!/bin/bash
# round 2, call AA (1 GPU): interior-gap skipping for sparse volumes: bitwise tests, DRR / fuzz / configs suites, bone-masked C2 timing,
# and the dense C2 bench with the gaps forced on (overhead of the map look-ups)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_skip_empty.py tests/test_gpu_drr.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/pytest_aa.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_aa.log
SWEEP_CASES=c2bone timeout 900 python scripts/sweep_configs.py > gpurun_out/r02_sweep_c2bone.jsonl 2> gpurun_out/r02_sweep_c2bone.err; cat gpurun_out/r02_sweep_c2bone.jsonl | cut -c1-400; tail -3 gpurun_out/r02_sweep_c2bone.err

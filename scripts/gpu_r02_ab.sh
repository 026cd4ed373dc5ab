#!/bin/bash
# round 2, call AB (1 GPU): final check of the whole GPU suite, smoke, default bench; then a scaled-up fuzz run with its counts
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c2_final2.log 2>&1
tail -c 400 gpurun_out/r02_bench_c2_final2.log
XREG_FUZZ_SCALE=6 timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02_fuzz_scale6.log 2>&1; echo "fuzz exit $?" >> gpurun_out/r02_fuzz_scale6.log
tail -4 gpurun_out/r02_fuzz_scale6.log

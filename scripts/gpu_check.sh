#!/bin/bash
# One gpurun call: GPU parity tests, smoke, C2 bench (own arm), latency script.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
tail -2 gpurun_out/bench_c2.log
timeout 300 python scripts/latency.py > gpurun_out/latency.log 2>&1
tail -20 gpurun_out/latency.log

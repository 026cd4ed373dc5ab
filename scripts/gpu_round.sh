#!/bin/bash
# One gpurun call: parity tests, smoke, bench, layout sweep.  Logs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --workload small --steps 5 --warmup 3 > gpurun_out/bench_small.log 2>&1
tail -2 gpurun_out/bench_small.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
tail -2 gpurun_out/bench_c2.log
timeout 900 python scripts/sweep_layouts.py > gpurun_out/sweep.log 2>&1
tail -40 gpurun_out/sweep.log

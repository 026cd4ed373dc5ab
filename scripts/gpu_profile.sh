#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# every launch of the bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -3 gpurun_out/launches_bench.log
# top kernel, full set, with source (AP view = bench default)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_kernel -s 4 -c 1 -f -o gpurun_out/prof_drr_ap \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_drr_ap.log 2>&1
tail -3 gpurun_out/prof_drr_ap.log
# metric kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|grad_kernel" -s 8 -c 2 -f -o gpurun_out/prof_sim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sim.log 2>&1
# lateral view (worst case of the sweep)
SWEEP_ONLY=quad,lateral,fine,0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_drr_lateral python scripts/sweep_layouts.py > gpurun_out/prof_drr_lateral.log 2>&1
ls -la gpurun_out/

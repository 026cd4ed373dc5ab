#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list + full captures of the top kernels.
# Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
# the first DRR launches are the fixed-image render and the count-only instrumentation passes: skip them
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_pax -s 10 -c 1 -f -o gpurun_out/prof_drr_trim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_drr_trim.log 2>&1
tail -2 gpurun_out/prof_drr_trim.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|grad_fast_kernel" -s 6 -c 2 -f -o gpurun_out/prof_sim2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sim2.log 2>&1
tail -2 gpurun_out/prof_sim2.log

#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
[ -x scripts/ubench/l1gather ] && timeout 120 scripts/ubench/l1gather > gpurun_out/l1gather.log 2>&1
# the bench itself, un-profiled (the numbers that count)
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
tail -1 gpurun_out/bench_c2.log
# every launch of the bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -3 gpurun_out/launches_bench.log
# top kernel, full set, with source (AP view = bench default)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_ -s 4 -c 1 -f -o gpurun_out/prof_drr_ap \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_drr_ap.log 2>&1
tail -3 gpurun_out/prof_drr_ap.log
# metric kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|grad_kernel" -s 8 -c 2 -f -o gpurun_out/prof_sim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sim.log 2>&1
ls -la gpurun_out/

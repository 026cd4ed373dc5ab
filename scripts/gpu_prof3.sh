#!/bin/bash
# ncu full captures of the bench kernels (after the count-only instrumentation launches) + configuration sweep.
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_pax -s 10 -c 1 -f -o gpurun_out/prof_drr_trim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_drr_trim.log 2>&1
tail -2 gpurun_out/prof_drr_trim.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|grad_fast_kernel" -s 6 -c 2 -f -o gpurun_out/prof_sim2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sim2.log 2>&1
tail -2 gpurun_out/prof_sim2.log
timeout 1500 python scripts/sweep_configs.py > gpurun_out/sweep_configs.log 2>&1
cat gpurun_out/sweep_configs.log

#!/bin/bash
# round 2, call M (1 GPU): ncu source-level capture of patch_seqsum2_kernel inside the 13-pose bench step
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_seqsum2 -s 2 -c 1 -f -o gpurun_out/prof_seqsum2 \
    python bench.py --batch 13 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_seqsum2.log 2>&1
tail -2 gpurun_out/prof_seqsum2.log
ls -la gpurun_out/prof_seqsum2.ncu-rep

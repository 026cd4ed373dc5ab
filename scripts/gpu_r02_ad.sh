#!/bin/bash
# round 2, call AD (1 GPU): depth ray caster timing on the C2 scene
set -x
mkdir -p gpurun_out
timeout 600 python scripts/depth_bench.py > gpurun_out/r02_depth_bench.jsonl 2> gpurun_out/r02_depth_bench.err; cat gpurun_out/r02_depth_bench.jsonl; tail -3 gpurun_out/r02_depth_bench.err

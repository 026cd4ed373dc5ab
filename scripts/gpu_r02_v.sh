#!/bin/bash
# round 2, call V (1 GPU): compute-sanitizer over the kernels that are new or changed in round 2 (patch_seqsum2 cluster kernel,
# xchg_kernel + tile-sharded stores on one rank, patch_kernel variants, on-demand stacks)
set -x
mkdir -p gpurun_out
SEL="tests/test_gpu_metrics.py tests/test_gpu_sharded_device.py tests/test_gpu_pipeline.py"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
      python -m pytest $SEL -m gpu -q -x -k "not full_size and not c2_population and not 480 and not 1600000 and not 557000 and not two_gpus" > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/r02_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Uninitialized|passed|failed|exit" gpurun_out/r02_sanitize_$tool.log | tail -6
done
timeout 300 python scripts/latency.py > gpurun_out/r02_latency.jsonl 2> gpurun_out/r02_latency.err; cat gpurun_out/r02_latency.jsonl | cut -c1-250

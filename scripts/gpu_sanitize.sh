#!/bin/bash
# compute-sanitizer over the GPU tests: memcheck on the whole suite, racecheck + initcheck on the small-size tests
set -x
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -q -x > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|exit" gpurun_out/sanitize_memcheck.log | tail -6
for tool in racecheck initcheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
      python -m pytest tests/test_gpu_drr.py tests/test_gpu_skip_empty.py tests/test_gpu_metrics.py tests/test_gpu_pipeline.py \
      -m gpu -q -x -k "not full_size and not c2_population and not 480" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Uninitialized|passed|failed|exit" gpurun_out/sanitize_$tool.log | tail -8
done

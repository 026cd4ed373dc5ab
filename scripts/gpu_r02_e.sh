#!/bin/bash
# round 2, call E: shared-memory staging microbenchmark (scripts/ubench/smem_stage.cu)
set -x
mkdir -p gpurun_out
B=scripts/ubench/smem_stage
: > gpurun_out/r02_smem_stage.jsonl
for args in "0" "1 0.40 1.25 4 3" "1 0.40 1.25 4 4" "1 0.40 1.25 8 2" "1 0.40 1.25 8 3" "1 0.40 1.25 2 4" \
            "0 0.40 1.25 4 3 13" "1 0.40 1.25 4 3 13" "0 0.40 1.25 4 3 100 0.0" "1 0.40 1.25 4 3 100 0.0" "0 0.40 1.25 4 3 100 0.15" "1 0.40 1.25 4 3 100 0.15" \
            "0 0.30 0.6 4 3" "1 0.30 0.6 4 3"; do
  timeout 120 $B $args >> gpurun_out/r02_smem_stage.jsonl 2>&1 || echo "{\"args\": \"$args\", \"failed\": $?}" >> gpurun_out/r02_smem_stage.jsonl
done
cat gpurun_out/r02_smem_stage.jsonl

#!/bin/bash
# round 2, call G: population-1 latency variants (deeper load pipeline, compact layouts) + the GPU suite
set -x
mkdir -p gpurun_out
: > gpurun_out/r02_latency_variants.jsonl
for det in 192 768; do
  LAT_ONLY=$det,grad-ncc,1,300 python scripts/latency.py >> gpurun_out/r02_latency_variants.jsonl 2>&1
  XRC_PAX_DEEP=8 LAT_ONLY=$det,grad-ncc,1,300 python scripts/latency.py >> gpurun_out/r02_latency_variants.jsonl 2>&1
  LAT_LAYOUT=linear LAT_ONLY=$det,grad-ncc,1,300 python scripts/latency.py >> gpurun_out/r02_latency_variants.jsonl 2>&1
  LAT_LAYOUT=quad LAT_ONLY=$det,grad-ncc,1,300 python scripts/latency.py >> gpurun_out/r02_latency_variants.jsonl 2>&1
done
cat gpurun_out/r02_latency_variants.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
